"""B200: fused InfoNCE pieces, device-side augmentation, the npz contract (same cases as the emulator tests, through the C ABI)."""
import pytest

import extra_cases as EC

pytestmark = pytest.mark.gpu


def test_infonce_kernels():
    EC.run_infonce_case("cuda")
    EC.run_infonce_case("cuda", B=1500, Pd=8, tau=0.1)


def test_augment_kernel():
    EC.run_augment_case("cuda")
    EC.run_augment_case("cuda", B=512, L=982, copies=3)


def test_npz_contract_and_augmenter():
    EC.run_npz_contract_case("cuda")

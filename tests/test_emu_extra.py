"""CPU (emulator build of csrc/extra.cu): fused InfoNCE pieces, device-side augmentation, the npz contract."""
import extra_cases as EC


def test_infonce_kernels(emu):
    EC.run_infonce_case("cpu")
    EC.run_infonce_case("cpu", B=70, Pd=8, tau=0.07)


def test_augment_kernel(emu):
    EC.run_augment_case("cpu")


def test_npz_contract_and_augmenter(emu):
    EC.run_npz_contract_case("cpu")

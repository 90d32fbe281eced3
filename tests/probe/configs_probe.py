import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch, bench
print(json.dumps(bench.run_configs(torch.device("cuda")), indent=1))

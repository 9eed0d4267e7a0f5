"""FP32-pipe probes (suhpe_fp32_probe): python profiles/probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json
import bench
from semiuhpe_b200 import _capi
dev = torch.device("cuda:0")
print(json.dumps(bench.fp32_probe(torch, _capi.lib(), dev, _capi), indent=1))

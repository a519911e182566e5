import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, json, bench
print(json.dumps(bench.c1_probe(torch.device('cuda:0'), nwarm=2, nrep=1)))

import sys, time, json
sys.path.insert(0, '.')
import torch
import bench
from visual_sgraphs_b200._lib import load
print(json.dumps(bench.other_config_extras(torch, load(), 0), indent=1))

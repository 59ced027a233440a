import sys, json; sys.path.insert(0,'/root/repo')
import torch, bench
from gaussianimage_plus_b200 import synth
print(json.dumps(bench.bench_fit_loop(torch, synth, 'cuda:0')))

"""Integer-pipe roofline microbenchmark: register-resident POPC / LOP3 / IADD /
engine-mix loops on the current GPU.  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kit4b_b200 as k4b

k4b.gpu_init(1)
names = {0: "popc", 1: "lop3", 2: "mix_2lop3_popc_min", 3: "iadd", 4: "imad", 5: "lop3_imad_1to1", 6: "shf",
         7: "imad_wide", 8: "lop3_imad_3to1", 9: "imad_hi", 10: "lop3_imad_hi_3to1"}
res = {}
for which, nm in names.items():
    res[nm + "_gops"] = round(k4b.microbench_intpipe(which, 4000), 1)
print(json.dumps(res))

import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],2), [(k, round(v["ms_per_launch"],3)) for k,v in d["kernels"].items() if v["ms_per_launch"]>0.25])

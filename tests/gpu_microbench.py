import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.Context(0, pkg.VESTA)
names = ["imad_wide", "imad32", "iadd3x", "madc_pairs", "fe_mul", "fe_addsub", "mul29_cols", "mul29_cios", "mul29_v3", "k9", "k10", "k11", "k12"]
res = {}
for kind, name in enumerate(names):
    try:
        res[name] = round(ctx.microbench(kind, 512), 1)
    except Exception as e:
        pass
print("MICROBENCH G-ops/s:", json.dumps(res), flush=True)

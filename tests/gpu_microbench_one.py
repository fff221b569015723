import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.Context(0, pkg.VESTA)
for kind in [int(x) for x in sys.argv[1].split(",")]:
    print(kind, ctx.microbench(kind, 64))

"""diagnostic: batched trp_dev_coeff_to_coset / lagrange_to_coeff over many columns at k = 20 against per-column calls"""
import sys
import torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.Context(0, pkg.VESTA)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 711
n = 1 << k
dom = pkg.EvaluationDomain(ctx, 6, k)
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.empty((cols, n, 4), dtype=torch.int64, device="cuda"); a.random_(0, 1 << 62, generator=g)
out = torch.empty_like(a)
one = torch.empty((n, 4), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for cs in (0, 4):
    ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, a.data_ptr(), out.data_ptr(), cols, cs)); ctx.sync()
    bad = []
    for c in list(range(0, cols, 37)) + [cols - 1, cols - 2, 127, 128, 255, 256, 511, 512]:
        if c >= cols: continue
        ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, a[c].data_ptr(), one.data_ptr(), 1, cs)); ctx.sync()
        if not torch.equal(one, out[c]): bad.append(c)
    print("coset", cs, "mismatching columns:", bad)

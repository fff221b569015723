"""Ad-hoc (not a test): one launch sequence of each hot kernel at benchmark size, for `ncu --set full -k regex:...` captures
(profiles/ncu_*_r02.md): 8 x (2^20 + 1) MSM (msm_accum_l1_seg_kernel), and one real proof at k = K (quotient_vm_kernel on a coset of
2^K rows with the real TinyRAM program; ntt passes).  usage: python tests/gpu_profile_kernels.py [msm|proof] [K]"""
import ctypes, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
what = sys.argv[1] if len(sys.argv) > 1 else "msm"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = pkg.Context(0, pkg.VESTA)
if what == "msm":
    from tiny_ram_halo2_b200 import synthetic
    ctx.bind_torch_stream()
    n, m = (1 << K) + 1, 8
    d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    synthetic.device_points(ctx, n, d_pts.data_ptr())
    hb = ctypes.c_void_p()
    ctx.check(ctx.lib.trp_dev_bases_load(ctx.handle, d_pts.data_ptr(), n, ctypes.byref(hb)))
    sc = torch.from_numpy(synthetic.random_scalars(n, 20, m).view(np.int64)).cuda()
    out = torch.zeros((m, 12), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, hb, sc.data_ptr(), n, m, out.data_ptr()))
    torch.cuda.synchronize()
else:
    from tiny_ram_halo2_b200 import plonk as PL, programs, tinyram as TR
    from tiny_ram_halo2_b200.sharded_backend import ShardedRng
    W = 32 if K >= 17 else 2 * (K - 2)
    tr = programs.longest_loop(W)
    circ, fixed, copies, adv, inst = TR.build(PL, tr, K, dense=False, arrays=True)
    be = PL.GpuBackend(ctx, K, circ.cs.degree())
    d_fixed, d_adv, d_inst = TR.device_columns(be, fixed), TR.device_columns(be, adv), TR.device_columns(be, inst)
    pk = PL.keygen(be, circ.cs, d_fixed, copies)
    rng = ShardedRng(be.p, seed=b"\x03" * 32)
    for _ in range(2):
        PL.create_proof(be, pk, d_inst, d_adv, rng, PL.Blake2bWrite(be.q, be.p))
    torch.cuda.synchronize()
print("done")

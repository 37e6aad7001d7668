"""Device-side timing of one conv layer shape: a recorded plan of `n` chained launches (ping-pong over 3 activation
buffers, like consecutive layers of the UNet), timed with one CUDA-event pair -> no host overhead in the number.
usage: layer_bench.py B H W Cin Cout [gn=1] [residual=1] [stats=1] [act=0] [taps=9] [n=24] [mod=0]
Set DDIF_LIB=<variant .so> to time a tuning build (csrc/build.py -D... -o...)."""
import sys, ctypes, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
from dif_pan_b200 import _lib


def bench(B, H, W, Cin, Cout, gn=1, res=1, st=1, act=0, taps=9, n=24, quiet=False, mod=0):
    g = torch.Generator().manual_seed(1)
    C = max(Cin, Cout)
    bufs = [torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(DEV) for _ in range(3)]
    k = 3 if taps == 9 else 1
    wp = pack_w((torch.randn(Cout, Cin, k, k, generator=g) * 0.03).to(DEV))
    bias = torch.zeros(Cout, device=DEV)
    gamma, beta = torch.ones(Cin, device=DEV), torch.zeros(Cin, device=DEV)
    stats_in = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    stats_in[:, 0] = 0.0
    stats_in[:, 1] = float(Cin * H * W)
    stats_out = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    modt = (torch.randn(B, H, W, 2 * Cout, generator=g) * 0.1).to(torch.bfloat16).to(DEV) if mod else None
    lib = _lib.load()
    plan = ctypes.c_void_p(lib.ddif_plan_create())
    for i in range(n):
        a, o, r = bufs[i % 3], bufs[(i + 1) % 3], bufs[(i + 2) % 3]
        p = _lib.make("ddif_gemm_t", a=[a.data_ptr(), 0], a_ld=[C, 0], a_c=[Cin, 0], a_h=[H, 0], a_w=[W, 0], w=[wp.data_ptr(), 0],
                      w_s=[wp.shape[0], 0], w_k=[wp.shape[2], 0], taps=[taps, 0], w_per_sample=[0, 0], nseg=1, stride=1, batch=B, out_h=H, out_w=W,
                      n_pad=(Cout + 15) // 16 * 16, n_valid=Cout, bias=bias.data_ptr(), film=None, film_ld=0, mod=modt.data_ptr() if mod else None,
                      residual=r.data_ptr() if res else None, res_ld=C if res else 0, act=act, out=o.data_ptr(), out_ld=C, out_nchw=None,
                      stats=stats_out.data_ptr() if st else None, gn_stats=stats_in.data_ptr() if gn else None,
                      gn_gamma=gamma.data_ptr() if gn else None, gn_beta=beta.data_ptr() if gn else None, gn_eps=1e-5, gn_act=1 if gn else 0,
                      a_up=0, force_tma=0, gn_stats2=None)
        rc = lib.ddif_plan_add(plan, _lib.KINDS["DDIF_OP_GEMM"], ctypes.byref(p))
        assert rc >= 0, rc
    s = stream()
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ddif_plan_run(plan, 0, -1, ctypes.c_void_p(s)), "run")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / n)
    lib.ddif_plan_destroy(plan)
    tiles = B * ((H + 15) // 16) * ((W + 7) // 8)
    fl = 2.0 * B * H * W * Cin * taps * Cout
    by = B * H * W * 2 * (Cin + Cout * ((2 if res else 1) + (2 if mod else 0)))
    if not quiet:
        print(f"B={B} {H}x{W} {Cin}->{Cout} k{taps} gn={gn} res={res} stats={st} act={act} mod={mod}: {best:7.1f} us/launch  "
              f"{best * 1965.0 / (tiles / 148.0):6.0f} cyc/tile  {fl / best / 1e6:6.0f} TFLOP/s  {by / best / 1e3:6.0f} GB/s", flush=True)
    return best


SUITE = [
    (256, 64, 64, 32, 32, 1, 1, 1, 0), (256, 64, 64, 32, 32, 1, 0, 1, 0), (256, 64, 64, 32, 32, 0, 0, 0, 0),
    (256, 32, 32, 64, 64, 1, 1, 1, 0), (256, 32, 32, 64, 64, 1, 0, 1, 0),
    (256, 16, 16, 64, 64, 1, 1, 1, 0), (256, 64, 64, 32, 64, 0, 0, 0, 1), (256, 64, 64, 64, 32, 0, 1, 1, 0),
    (256, 32, 32, 64, 128, 0, 0, 0, 1), (256, 32, 32, 128, 64, 0, 1, 1, 0), (256, 64, 64, 64, 64, 1, 0, 0, 0),
    (256, 32, 32, 128, 128, 1, 0, 0, 0),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "suite":
        a = [int(v) for v in sys.argv[1:]]
        d = [1, 1, 1, 0, 9, 24, 0]
        a = a + d[len(a) - 5:]
        bench(*a[:5], gn=a[5], res=a[6], st=a[7], act=a[8], taps=a[9], n=a[10], mod=a[11])
    else:
        for s_ in SUITE:
            bench(*s_)

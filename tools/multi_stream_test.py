"""Experiment: one UNet denoise step of B patches as S concurrent sub-batches (one recorded plan + CUDA graph + stream
each) vs one plan over the whole batch.  Persistent kernels of different sub-batches fill each other's tails/prologues
and the small low-resolution grids share the GPU.
usage: multi_stream_test.py [B=256] [S=1,2,4]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dif_pan_b200 as dp
from dif_pan_b200 import synth
from dif_pan_b200.unet import _Runtime

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
kw = synth.unet_kwargs("wv3")
net = dp.UNetSR3(**kw)
net.load_state_dict(synth.make_state_dict(0, **kw))
net = net.to(dev).eval()
net.pack_weights()
cond = synth.make_batch("wv3", 16, seed=1)["cond"]
cond = cond.repeat((B + 15) // 16, 1, 1, 1)[:B].contiguous().to(dev)
for S in SS:
    b = B // S
    rts = [_Runtime(net, b, 64, 64) for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    for i, (rt, st) in enumerate(zip(rts, streams)):
        with torch.cuda.stream(st):
            rt.set_cond(cond[i * b:(i + 1) * b].contiguous())
            dp.diffusion.device_randn_(rt.x_buf, 1, i)
            rt.t_buf.fill_(250.0)
            rt.step()
    torch.cuda.synchronize()
    main = torch.cuda.current_stream(dev)

    def step_all():
        for rt, st in zip(rts, streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                rt.step()
        for st in streams:
            main.wait_stream(st)

    for _ in range(3):
        step_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        step_all()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} as {S} x {b}: {e0.elapsed_time(e1) / reps:.3f} ms per denoise step", flush=True)
    for rt in rts:
        rt.close()
    del rts
    torch.cuda.empty_cache()

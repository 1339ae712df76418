"""tools/slab_overhead.py -- where the non-kernel time of a slab build goes (torchrun, N ranks): CUDA-event times of
the summary kernel, the all-gather and the slab build, per rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sdf_tools_b200 import Context, synth
from sdf_tools_b200.distributed import SlabBuilder

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
lr = int(os.environ.get("LOCAL_RANK", "0")); rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = synth.CONFIGS[wl]; nx, ny, nz = cfg["shape"]
ctx = Context(lr); sb = SlabBuilder(ctx, (nx, ny, nz), rank, world, parts=1)
mask = torch.from_numpy(synth.make_mask(wl, zrange=(sb.z0, sb.z1))).cuda()
out = torch.empty((nx, ny, sb.nzl), dtype=torch.float32, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    ctx.slab_summary_device(mask.data_ptr(), nx, ny, sb.nzl, sb.summary.data_ptr())
    ev[1].record()
    dist.all_gather_into_tensor(sb.all_summaries, sb.summary)
    ev[2].record()
    ctx.slab_build_device(mask.data_ptr(), nx, ny, sb.nzl, cfg["res"], sb.all_summaries.data_ptr(), sb.widths, rank, out.data_ptr())
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 3 and rank == 0:
        print("summary %.3f ms  all_gather %.3f ms  carry+3 passes %.3f ms  (kernels %s)" % (
            ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ctx.last_timing()))
dist.destroy_process_group()

"""N-GPU check of the build-once / broadcast path. Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/replicate_check.py [B|C|D]
Rank 0 builds the index, exports it, NCCL broadcasts the image over NVLink, the other ranks import it.
Every rank then joins its query shard on the received index AND on an index it built itself: offsets and
pairs must be identical. Prints the time of both ways of getting an index onto every GPU."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from binary_b200 import DeviceIndex, synth
from binary_b200.sharding import replicate_index, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "B"]
tg, tl, th = w.targets()
n_q = 2_000_000
start, stop = shard_range(n_q * world, rank, world)
qg, ql, qh = w.queries(start, stop - start)
t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
d_tg, d_tl, d_th = map(t, (tg, tl, th))
stream = torch.cuda.current_stream().cuda_stream
build = lambda: DeviceIndex.build_dev(tl.size, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr(), device=local, stream=stream)

def timed(fn, reps=5):
    out, best = None, 1e9
    for _ in range(reps):
        if out is not None: out.close()
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(); dist.barrier(); best = min(best, time.perf_counter() - t0)
    return out, best

mine, t_build = timed(build)
shipped, t_bcast = timed(lambda: replicate_index(build, local, src=0))
a = mine.join(ql, qh, qg)
b = shipped.join(ql, qh, qg)
same = all(np.array_equal(x, y) for x, y in zip(a, b)) and mine.info() == shipped.info()
flags = torch.tensor([int(same)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{w.name}: world {world}, image {shipped.image_size()/1e6:.1f} MB, every rank builds: {t_build*1e3:.2f} ms, "
          f"build once + NCCL broadcast + import: {t_bcast*1e3:.2f} ms, identical results on all ranks: {bool(flags.item())}",
          flush=True)
dist.destroy_process_group()
sys.exit(0 if flags.item() else 1)

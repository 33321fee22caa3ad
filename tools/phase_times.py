"""Per-phase CUDA-event times of one hot-path step on every rank (run under torch.distributed.run for N > 1).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/phase_times.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from stmask_b200 import sharding
from stmask_b200.hotpath import HotPath, HotPathConfig

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else ("frame" if world > 1 else "clip")
plan = sharding.make_plan(2 * world, 36, world, mode)
hp = HotPath(HotPathConfig(), dev)
inp = hp.make_inputs(plan.local_frames(rank), dev, seed=rank)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e = [ev() for _ in range(5)]
    e[0].record()
    halo = None
    if world > 1 and plan.recv_halos(rank) or plan.send_halos(rank):
        halo = sharding.exchange_halo(plan, rank, [inp["tf.fpn"], inp["tf.t2s"]], None)
    e[1].record()
    hp._frames_only({k: v for k, v in inp.items() if not k.startswith("tf.")})
    e[2].record()
    fr, fn = sharding.temporal_pairs(plan, rank, inp["tf.fpn"], halo[0] if halo else None)
    tr, tn = sharding.temporal_pairs(plan, rank, inp["tf.t2s"], halo[1] if halo else None)
    e[3].record()
    hp.temporal_fusion(fr, fn, tr, tn)
    e[4].record()
    torch.cuda.synchronize()
    e5 = [ev(), ev()]
    e5[0].record(); hp(inp, plan, rank); e5[1].record(); torch.cuda.synchronize()
    if it >= 3:
        print(f"rank {rank} it {it} [{mode}]: halo {e[0].elapsed_time(e[1]):.3f}  dcn+fcb {e[1].elapsed_time(e[2]):.3f}  pairs {e[2].elapsed_time(e[3]):.3f}  "
              f"tf {e[3].elapsed_time(e[4]):.3f}  | whole forward {e5[0].elapsed_time(e5[1]):.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()

"""Final NVLink gather of encoder states on real NCCL: correctness on a sharded tiny model, then the time of the gather
at the bench size (32 chunks x 1500 x 1280 per rank, bf16 and f32).  Launch: torchrun --nproc-per-node N tools/gather_check.py"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_apr_b200 import WhisperApr, sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = synth.CONFIGS["tiny"]
data, _ = synth.random_model_apr(cfg, seed=0)
model = WhisperApr.load_from_apr(data, device=local)
n_chunks = 5
s, e = sharding.shard_range(n_chunks, world, rank)
mine = [synth.synth_audio(100 + c) for c in range(s, e)]
states = torch.from_numpy(model.mel_encode_batch(mine)).cuda() if mine else torch.zeros((0, 1500, cfg.n_audio_state), device="cuda")
full = sharding.gather_states(states, n_chunks)
ref = model.mel_encode_batch([synth.synth_audio(100 + c) for c in range(n_chunks)])      # every rank recomputes all chunks locally
ok = bool(np.array_equal(full.cpu().numpy(), ref))
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
    x = torch.randn((32, 1500, 1280), device="cuda").to(dt)
    for _ in range(3):
        sharding.gather_states(x, 32 * world)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sharding.gather_states(x, 32 * world)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes = x.numel() * x.element_size()
        print(f"gather {name}: {ms.item():.3f} ms per call for {nbytes / 1e6:.0f} MB per rank x {world} ranks "
              f"({nbytes * (world - 1) / ms.item() / 1e6:.0f} GB/s received per GPU)", flush=True)
if rank == 0:
    print("gathered states identical to the single-rank result on every rank:", bool(flag.item() == 1.0), flush=True)
model.close()
dist.destroy_process_group()

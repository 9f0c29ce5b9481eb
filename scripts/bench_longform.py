#!/usr/bin/env python
"""BASELINE.json configs[3]: long-form 10-minute synthetic audio, chunked with overlap, streamed through encoder + VQ
(bench_configs.longform).  1 GPU: python scripts/bench_longform.py     N GPUs: torchrun --nproc-per-node N ... scripts/bench_longform.py
The conv front end is dealt out over the ranks chunk by chunk; the frame-rate features make one ordered hand-off to rank 0,
which runs the LSTM + final conv + VQ over the whole recording."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench_configs
from audiotokenization_b200 import configs, synth
from audiotokenization_b200.model import BigCodecModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--chunk-seconds", type=float, default=30.0)
    ap.add_argument("--micro-batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check-whole", action="store_true", help="also encode the recording in one piece and compare")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = configs.get_config("base")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
    model = BigCodecModel(cfg, enc_sd, dec_sd, device=f"cuda:{local}", precision=args.precision)
    line = bench_configs.longform(model, args.minutes, args.chunk_seconds, args.micro_batch, args.steps, world, args.check_whole)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

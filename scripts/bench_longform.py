#!/usr/bin/env python
"""BASELINE.json configs[3]: long-form 10-minute synthetic audio, chunked with overlap, streamed through encoder + VQ.
1 GPU: python scripts/bench_longform.py     N GPUs: torchrun --nproc-per-node N ... scripts/bench_longform.py
The conv front end is dealt out over the ranks chunk by chunk; the frame-rate features make one ordered hand-off
(NCCL gather) to rank 0, which runs the LSTM + final conv + VQ over the whole recording."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
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
    hop = int(model.encoder.hop_length)
    T = int(args.minutes * 60 * 16000) // hop * hop
    x = synth.fast_synth_batch(99, 1, T)[0, 0].cuda()

    def step():
        return model.indices_longform(x, chunk_seconds=args.chunk_seconds, micro_batch=args.micro_batch)

    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"workload": f"configs[3] long-form: one {args.minutes:g}-minute recording, {args.chunk_seconds:g} s chunks + halo, "
                            f"front end over {world} GPU(s), LSTM + VQ on rank 0", "precision": args.precision, "n_gpus": world,
                "ms_per_recording": float(ms), "audio_s_per_s": args.minutes * 60 / float(ms) * 1e3, "frames": int(out.shape[1])}
        if args.check_whole:
            whole = model.indices_device(x.view(1, 1, -1), micro_batch=1, rnn_batch=1)
            line["equals_unchunked"] = bool(torch.equal(whole, out))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

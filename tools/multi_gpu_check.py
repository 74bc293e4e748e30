#!/usr/bin/env python
"""Multi-GPU check of the sharded overall-style statistics, run under
torchrun on the GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [--images 64] [--size 128] [--batch 8]

Every rank encodes its contiguous share of one synthetic client (mean_std_computation_effcientMem.py
:117-131), the per-rank Welford states meet in ONE NCCL all-reduce of the 1+2C fp64 moments, and
rank 0 compares the result with a single-GPU pass over all images (same kernels, no collective) and,
for small sizes, with the fp64 oracle of the reference formula.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ccst_b200
from ccst_b200 import overall, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--oracle", action="store_true")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vgg, dec = synth.make_models(0)
    eng = ccst_b200.engine_for(vgg, dec, dev)

    def batch_of(first, count):  # image i of the client has seed 5000 + i on every rank
        return torch.cat([synth.images(1, a.size, a.size, 5000 + i) for i in range(first, first + count)])

    def run(begin, end, group_reduce):
        acc = overall.OverallStyleAccumulator(eng, a.precision)
        for b0 in range(begin, end, a.batch):
            acc.add_images(batch_of(b0, min(a.batch, end - b0)).to(dev))
        if group_reduce:
            return acc.finalize(None) + (acc.global_img_count,)
        return acc.state.finalize() + (acc.img_count,)

    begin, end = overall.shard_range(a.images, rank, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean, std, seen = run(begin, end, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = {"world": world, "images": a.images, "size": a.size, "precision": a.precision, "seen": seen,
           "sharded_s": round(dt, 3)}

    # the same merge through the C ABI alone (ccst_allreduce_moments on a raw ncclComm_t): what a host without
    # torch.distributed does; torch.distributed is used here only to ship the 128-byte NCCL id to the ranks
    from ccst_b200 import _lib, nccl_raw
    from ccst_b200 import function as F_

    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_raw.unique_id()), dtype=torch.uint8).clone()
    if world > 1:
        uid_d = uid.to(dev)
        dist.broadcast(uid_d, 0)
        uid = uid_d.cpu()
    comm = nccl_raw.comm_init(world, rank, bytes(uid.numpy().tobytes()))
    acc = overall.OverallStyleAccumulator(eng, a.precision)
    for b0 in range(begin, end, a.batch):
        acc.add_images(batch_of(b0, min(a.batch, end - b0)).to(dev))
    payload = torch.cat([acc.state.moments(), torch.tensor([float(acc.img_count)], dtype=torch.float64, device=dev)])
    st = torch.cuda.current_stream(dev)
    _lib.check(_lib.lib().ccst_allreduce_moments(comm, payload.data_ptr(), payload.numel(), st.cuda_stream))
    merged = F_.WelfordState(512, dev).load_moments(payload[:-1])
    mean_c, std_c = merged.finalize()
    torch.cuda.synchronize()
    res["c_abi_allreduce_equals_torch_distributed"] = bool(torch.equal(mean_c, mean) and torch.equal(std_c, std)
                                                           and int(round(payload[-1].item())) == seen)
    nccl_raw.comm_destroy(comm)
    if rank == 0:
        m1, s1, n1 = run(0, a.images, False)
        res["single_gpu_seen"] = n1
        res["mean_rel_vs_single"] = ((mean - m1).abs().max() / m1.abs().max()).item()
        res["std_rel_vs_single"] = ((std - s1).abs().max() / s1.abs().max()).item()
        if a.oracle:
            from oracle import ccst_oracle as O

            with torch.no_grad():
                feats = [O.encode_relu4_1(vgg, batch_of(b0, min(a.batch, a.images - b0)))
                         for b0 in range(0, a.images, a.batch)]
            m64, s64, _, _ = O.overall_style_stats(feats, dtype=torch.float64)
            res["mean_rel_vs_oracle64"] = ((mean.cpu().double() - m64).abs().max() / m64.abs().max()).item()
            res["std_rel_vs_oracle64"] = ((std.cpu().double() - s64).abs().max() / s64.abs().max()).item()
        ok = (seen == a.images and res["mean_rel_vs_single"] < 1e-5 and res["std_rel_vs_single"] < 1e-5
              and res["c_abi_allreduce_equals_torch_distributed"])
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

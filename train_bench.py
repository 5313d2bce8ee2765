#!/usr/bin/env python
"""train_bench.py — the second half of BASELINE.json's metric: transformer train tokens/s at
1/2/4/8 B200 (configs[3] and configs[4]), measured on the op streams burn-nn / burn-autodiff /
burn-optim issue (burn_b200/train.py) through the burn_b200 C ABI.

  python train_bench.py --config lm --steps 10 --warmup 3 [--mm tf32|bf16|f32x3] [--eager]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P train_bench.py --config lm ...

One step = upload this rank's token batch from pinned host memory, forward, loss, backward with
every parameter gradient all-reduced (Mean) over NCCL as soon as it is final (N > 1), Adam on every
parameter, loss read back to the host.  The device work is captured ONCE into a CUDA graph
(b200_graph_*) and replayed; `--eager` replays the Python launch sequence instead.  Weak scaling:
the per-GPU batch is fixed, tokens/s is the whole-job aggregate, time is max over ranks on the device.
bench.py imports `run()` and attaches its result to the bench line under "train".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# configs[3]: burn-nn TransformerEncoder d_model 512, 6 layers, 8 heads, seq 256, batch 64 (d_ff 2048: 4·d_model)
# configs[4]: text-generation-style LM d_model 1024, 12 layers, seq 1024 (16 heads, d_ff 4096; vocab = GPT-2's
#             50257 + [START]/[END]/[PAD] = 50260, examples/text-generation/src/data/tokenizer.rs:27-31);
#             per-GPU batch 8 → 8192 tokens per GPU per step (SURVEY.md §8 sizes "T5")
CONFIGS = {
    "encoder": dict(kind="encoder", d=512, ff=2048, h=8, L=6, S=256, B=64),
    "lm": dict(kind="lm", d=1024, ff=4096, h=16, L=12, S=1024, B=8, vocab=50260),
    "lm-tiny": dict(kind="lm", d=64, ff=128, h=4, L=2, S=32, B=2, vocab=96),
    "encoder-tiny": dict(kind="encoder", d=64, ff=128, h=4, L=2, S=16, B=4),
}


def model_flops(cfg) -> float:
    """fwd+bwd FLOPs per step per GPU: 6·(GEMM params)·tokens + 12·L·B·S²·d for attention."""
    d, ff, L, S, B = cfg["d"], cfg["ff"], cfg["L"], cfg["S"], cfg["B"]
    tokens = B * S
    p = L * (4 * d * d + 2 * d * ff)
    if cfg["kind"] == "lm":
        p += d * cfg["vocab"]
    return 6.0 * p * tokens + 12.0 * L * B * S * S * d


def run(cfg_name: str, steps: int, warmup: int, rank: int, world: int, local_rank: int,
        mm: str = "tf32", use_graph: bool = True, bucket_mb: int = 32, init_device: bool = True,
        sync: str | None = None):
    """sync: how gradients are synchronised at world > 1 —
         "nccl"   one ncclAllReduce(avg) per bucket on the collective stream, multi-tensor Adam afterwards
         "peer"   the peer-memory all-reduce kernel (burn_b200/csrc/peer.cu) per bucket, Adam afterwards
         "fused"  reduce-scatter → Adam on the owned 1/N shard → parameter all-gather in ONE peer-memory kernel per bucket
       default: $B200_GRAD_SYNC, else "fused"."""
    import torch
    import torch.distributed as dist
    from burn_b200 import _abi as abi
    from burn_b200 import device as dv
    from burn_b200 import train as T
    from burn_b200 import ops
    from burn_b200.device import DeviceTensor
    from burn_b200.distributed import Communicator

    cfg = CONFIGS[cfg_name]
    prec = {"tf32": abi.MM_TF32, "bf16": abi.MM_BF16, "f32x3": abi.MM_F32X3}[mm]
    if init_device:
        torch.cuda.set_device(local_rank)
        dv.init(local_rank)
    lib = abi.load()
    check = abi.check
    B, S, d = cfg["B"], cfg["S"], cfg["d"]

    # ---- model (same seed on every rank = identical replicas), optimizer, gradient sync
    if cfg["kind"] == "lm":
        model = T.LanguageModel(11, cfg["vocab"], S, d, cfg["ff"], cfg["h"], cfg["L"])
    else:
        model = T.Encoder(11, d, cfg["ff"], cfg["h"], cfg["L"])
    params = model.params()
    n_params = sum(p.v.numel for p in params)
    opt = T.Adam(lr=1e-4)
    comm = peer = None
    sync = (sync or os.environ.get("B200_GRAD_SYNC", "fused")) if world > 1 else "none"
    if sync not in ("none", "nccl", "peer", "fused"):
        raise ValueError(f"unknown gradient sync mode {sync!r}")
    if sync == "nccl":
        comm = Communicator(rank, world, device=torch.device("cuda", local_rank))
    elif sync in ("peer", "fused"):
        from burn_b200.distributed import PeerGroup
        peer = PeerGroup(rank, world, T.ParamArena.peer_bytes(params), device=torch.device("cuda", local_rank))
    # flat p/m/v/g buckets: direct-to-bucket gradients, one all-reduce and one Adam launch per bucket
    arena = T.ParamArena(params, comm, bucket_bytes=bucket_mb << 20, peer=peer, fused=sync == "fused")
    arena.attach_optimizer(opt)

    # ---- this rank's synthetic batch, pinned on the host, persistent device buffers
    rng = np.random.default_rng(5000 + rank)

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        check(lib.b200_host_alloc(C.byref(ptr), n))
        ctype = {np.int32: C.c_int32, np.float32: C.c_float}[dtype]
        return ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape), n

    loss_dev = DeviceTensor.empty((1,))
    loss_ptr, loss_host, _ = pinned((1,), np.float32)
    if cfg["kind"] == "lm":
        tok_ptr, tok_host, tok_bytes = pinned((B, S), np.int32)
        tgt_ptr, tgt_host, _ = pinned((B, S), np.int32)
        tok_host[...] = rng.integers(0, cfg["vocab"], (B, S))
        tgt_host[...] = np.roll(tok_host, -1, axis=1)                    # next-token targets
        tok_dev, tgt_dev = DeviceTensor.empty((B, S), abi.I32), DeviceTensor.empty((B, S), abi.I32)
        pos_dev = DeviceTensor.from_numpy(np.tile(np.arange(S, dtype=np.int32), (B, 1)))
        causal = DeviceTensor.from_numpy(np.triu(np.ones((S, S), dtype=bool), k=1)[None, None])
        h2d_bytes = 2 * tok_bytes

        def h2d():
            check(lib.b200_memcpy_h2d(tok_dev.data_ptr(), tok_ptr, tok_bytes, None))
            check(lib.b200_memcpy_h2d(tgt_dev.data_ptr(), tgt_ptr, tok_bytes, None))
    else:
        x_ptr, x_host, x_bytes = pinned((B, S, d), np.float32)
        x_host[...] = rng.standard_normal((B, S, d)).astype(np.float32)
        x_dev = DeviceTensor.empty((B, S, d))
        h2d_bytes = x_bytes

        def h2d():
            check(lib.b200_memcpy_h2d(x_dev.data_ptr(), x_ptr, x_bytes, None))

    def device_step():
        """forward → loss → backward (+ overlapped all-reduce) → Adam; every buffer it allocates dies here."""
        tape = T.Tape(prec)
        if cfg["kind"] == "lm":
            loss = model.loss(tape, tok_dev, tgt_dev, pos_dev, causal)
        else:
            loss = T.mean_square(tape, model.forward(tape, T.Var(x_dev, False)))
        check(lib.b200_memcpy_d2d(loss_dev.data_ptr(), loss.v.data_ptr(), 4, None))
        del loss
        tape.backward()
        arena.wait()
        opt.apply_arena(arena)
        T.Adam.zero_grad(params)

    def d2h():
        check(lib.b200_memcpy_d2h(loss_ptr, loss_dev.data_ptr(), 4, None))

    def barrier():
        check(lib.b200_device_sync())
        if world > 1:
            dist.barrier()
        check(lib.b200_device_sync())

    # ---- one eager step (creates the Adam moments, sets kernel attributes), then capture
    losses = []
    h2d(); opt.advance(); device_step(); d2h()
    check(lib.b200_device_sync())
    losses.append(float(loss_host[0]))
    graph = None
    if use_graph:
        with dv.Graph.capture() as graph:
            device_step()

    def step():
        h2d()
        opt.advance()
        if graph is not None:
            graph.launch()
        else:
            device_step()
        d2h()

    for _ in range(warmup):
        step()
        check(lib.b200_device_sync())
        losses.append(float(loss_host[0]))
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(ev0)))
    check(lib.b200_event_create(C.byref(ev1)))
    barrier()
    lib.b200_launch_count_reset()
    check(lib.b200_event_record(ev0, None))
    for _ in range(steps):
        step()
    check(lib.b200_event_record(ev1, None))
    barrier()
    launches = int(lib.b200_launch_count())
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    total_ms = ms.value
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    losses.append(float(loss_host[0]))
    # replicas must still be identical after the run: compare a checksum of every parameter bucket across ranks
    replica_spread = None
    if world > 1:
        sums = [float(ops.float_sum(b["p"]).numpy()[0]) for b in arena.buckets]
        t = torch.tensor(sums, device="cuda", dtype=torch.float64)
        hi, lo = t.clone(), t.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        replica_spread = float((hi - lo).abs().max().item())
    ms_per_step = total_ms / steps
    tokens = B * S * world
    flops = model_flops(cfg) * world
    out = {
        "metric": "transformer train tokens/s", "value": round(tokens / (ms_per_step * 1e-3), 1), "unit": "tokens/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms_per_step, 3),
        "scaling": "weak", "dtype": {"tf32": "f32 storage, tf32 tensor-core GEMMs", "bf16": "f32 storage, bf16 GEMM operands, f32 accumulate",
                                      "f32x3": "f32 storage, 3xTF32 split GEMMs"}[mm],
        "config": {"workload": f"configs[{3 if cfg['kind'] == 'encoder' else 4}] " + (
            f"TransformerEncoder d_model {d}, {cfg['L']} layers, {cfg['h']} heads, seq {S}, batch {B}/GPU, fwd+bwd+Adam"
            if cfg["kind"] == "encoder" else
            f"decoder LM d_model {d}, {cfg['L']} layers, {cfg['h']} heads, d_ff {cfg['ff']}, vocab {cfg['vocab']}, seq {S}, "
            f"batch {B}/GPU, fwd+bwd+NCCL grad all-reduce+Adam"),
            "params": n_params, "tokens_per_step": tokens,
            "parallelism": f"dp{world}" + {
                "none": f"; multi-tensor Adam: {len(arena.buckets)} launches",
                "nccl": f", {len(arena.buckets)} ncclAllReduce(avg) per step on ~{bucket_mb} MiB flat buckets, overlapped with "
                        f"backward on the collective stream; multi-tensor Adam: {len(arena.buckets)} launches",
                "peer": f", {len(arena.buckets)} peer-memory all-reduce kernels (NVLink loads/stores, no NCCL) per step on "
                        f"~{bucket_mb} MiB flat buckets beside backward; multi-tensor Adam: {len(arena.buckets)} launches",
                "fused": f", {len(arena.buckets)} fused reduce-scatter + Adam(1/{world} shard) + parameter all-gather kernels over "
                         f"NVLink peer memory per step (~{bucket_mb} MiB flat buckets), running beside backward; optimizer state "
                         f"sharded {world} ways; no NCCL on the data path"}[sync],
            "grad_sync": sync,
            "launch": "cuda graph replay" if graph is not None else "eager (python launch loop)"},
        "model_tflops_per_s": round(flops / (ms_per_step * 1e-3) / 1e12, 1),
        "gpu_launches": launches, "kernels_per_step": launches // max(steps, 1),
        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
        "loss_first": round(losses[0], 5), "loss_last": round(losses[-1], 5),
    }
    if replica_spread is not None:
        out["replica_checksum_spread"] = replica_spread     # 0.0: every rank holds bit-identical parameters
    if graph is not None:
        check(lib.b200_device_sync())
        graph.destroy()             # before the communicator: NCCL waits for graphs that captured it
    if comm is not None:
        comm.close()
    if peer is not None:
        barrier()                   # nobody may still be writing into a region that is about to be unmapped
        peer.close()
    return out


def run_fc_head(steps: int = 200, warmup: int = 10, mm: str = "f32x3"):
    """BASELINE.json configs[0] on the GPU path: the MNIST example's FC head (burn_b200.train.FcHead), batch 64,
    one captured graph per step; every step uploads a fresh [64, 1600] batch from pinned memory and reads the loss."""
    from burn_b200 import _abi as abi
    from burn_b200 import device as dv
    from burn_b200 import train as T
    from burn_b200.device import DeviceTensor
    lib, check = abi.load(), abi.check
    prec = {"tf32": abi.MM_TF32, "bf16": abi.MM_BF16, "f32x3": abi.MM_F32X3}[mm]
    model = T.FcHead(0)
    params, opt = model.params(), T.Adam(lr=1e-3)
    arena = T.ParamArena(params, None)
    B, D = 64, T.FcHead.DIMS[0]
    rng = np.random.default_rng(9)
    xp, tp, lp = C.c_void_p(), C.c_void_p(), C.c_void_p()
    check(lib.b200_host_alloc(C.byref(xp), B * D * 4))
    check(lib.b200_host_alloc(C.byref(tp), B * 4))
    check(lib.b200_host_alloc(C.byref(lp), 4))
    xh = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_float)), shape=(B, D))
    th = np.ctypeslib.as_array(C.cast(tp, C.POINTER(C.c_int32)), shape=(B,))
    lh = np.ctypeslib.as_array(C.cast(lp, C.POINTER(C.c_float)), shape=(1,))
    xh[...] = np.maximum(rng.standard_normal((B, D)), 0).astype(np.float32)
    th[...] = rng.integers(0, 10, B)
    xd, td, ld = DeviceTensor.empty((B, D)), DeviceTensor.empty((B,), abi.I32), DeviceTensor.empty((1,))

    def device_step():
        tape = T.Tape(prec)
        loss = model.loss(tape, xd, td)
        check(lib.b200_memcpy_d2d(ld.data_ptr(), loss.v.data_ptr(), 4, None))
        del loss
        tape.backward()
        arena.wait()
        opt.apply_arena(arena)
        T.Adam.zero_grad(params)

    def io_in():
        check(lib.b200_memcpy_h2d(xd.data_ptr(), xp, B * D * 4, None))
        check(lib.b200_memcpy_h2d(td.data_ptr(), tp, B * 4, None))

    io_in(); opt.advance(); device_step()
    check(lib.b200_device_sync())
    with dv.Graph.capture() as graph:
        device_step()

    def step():
        io_in()
        opt.advance()
        graph.launch()
        check(lib.b200_memcpy_d2h(lp, ld.data_ptr(), 4, None))

    for _ in range(warmup):
        step()
    check(lib.b200_device_sync())
    first = float(lh[0])
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(e0))); check(lib.b200_event_create(C.byref(e1)))
    check(lib.b200_event_record(e0, None))
    for _ in range(steps):
        step()
    check(lib.b200_event_record(e1, None))
    check(lib.b200_device_sync())
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    graph.destroy()
    return {"workload": "configs[0]: examples/mnist FC head 1600-128-128-10, batch 64, fwd+bwd+Adam, one CUDA graph per step, "
                        "H2D of the batch and D2H of the loss inside the timed region",
            "value": round(B * steps / (ms.value * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(ms.value / steps, 4),
            "gemm": mm, "loss_first": round(first, 4), "loss_last": round(float(lh[0]), 4)}


def run_via_stream(steps: int = 20, warmup: int = 3, mm: str = "tf32"):
    """configs[3] encoder FORWARD issued as the primitive operation stream burn-nn emits (burn_b200/stream_model.py),
    planned and launched by the host fusion layer — beside the same forward sequenced by hand (burn_b200/train.py,
    the path `run` measures).  The stream materialises the attention weights (burn-nn's MultiHeadAttention does;
    the hand-sequenced forward is timed both ways: the same chain, and the flash kernel)."""
    import time
    from collections import Counter
    from burn_b200 import _abi as abi
    from burn_b200 import device as dv
    from burn_b200 import fusion as F
    from burn_b200 import train as T
    from burn_b200.device import DeviceTensor
    from burn_b200.stream_model import StreamEncoder
    lib, check = abi.load(), abi.check
    cfg = CONFIGS["encoder"]
    prec = {"tf32": abi.MM_TF32, "bf16": abi.MM_BF16, "f32x3": abi.MM_F32X3}[mm]
    dv.init(int(os.environ.get("LOCAL_RANK", "0")))
    B, S, d = cfg["B"], cfg["S"], cfg["d"]
    model = T.Encoder(11, d, cfg["ff"], cfg["h"], cfg["L"])
    x_dev = DeviceTensor.from_numpy(np.random.default_rng(3).standard_normal((B, S, d)).astype(np.float32))
    out_stream, out_hand = DeviceTensor.empty((B, S, d)), DeviceTensor.empty((B, S, d))
    nbytes = B * S * d * 4

    def timed(fn, n):
        e0, e1 = C.c_void_p(), C.c_void_p()
        check(lib.b200_event_create(C.byref(e0))); check(lib.b200_event_create(C.byref(e1)))
        check(lib.b200_device_sync())
        t0 = time.perf_counter()
        check(lib.b200_event_record(e0, None))
        for _ in range(n):
            fn()
        check(lib.b200_event_record(e1, None))
        host_ms = (time.perf_counter() - t0) * 1e3 / n
        check(lib.b200_device_sync())
        ms = C.c_float()
        check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value / n, host_ms

    # ---- hand-sequenced forward (attention as the configured mode, and as the unfused chain)
    def hand_forward():
        tape = T.Tape(prec)
        y = model.forward(tape, T.Var(x_dev, False))
        check(lib.b200_memcpy_d2d(out_hand.data_ptr(), y.v.data_ptr(), nbytes, None))
    res = {}
    for mode in ("chain", "flash"):
        T._ATTN_MODE = mode
        for _ in range(warmup):
            hand_forward()
        lib.b200_launch_count_reset()
        ms, host = timed(hand_forward, steps)
        res[f"hand_{mode}"] = {"ms": round(ms, 3), "host_ms": round(host, 3), "launches": int(lib.b200_launch_count()) // steps}
    T._ATTN_MODE = "chain"
    hand_forward()                       # reference values for the stream: same attention decomposition
    check(lib.b200_device_sync())
    want = out_hand.numpy()

    # ---- the same forward as an operation stream
    st = F.FusionStream()
    enc = StreamEncoder(st, model, prec)
    xs = st.wrap(x_dev.desc())

    def stream_forward():
        y = enc.forward(xs)
        d_ = y.device_tensor()            # flush: plan (or fetch the cached plan) and launch
        check(lib.b200_memcpy_d2d(out_stream.data_ptr(), d_.ptr, nbytes, None))
        y.drop()                          # queue is empty here: the handle (and its buffer) is released at once
    t0 = time.perf_counter()
    stream_forward()
    check(lib.b200_device_sync())
    first_ms = (time.perf_counter() - t0) * 1e3
    blocks = st.blocks()
    kinds = Counter({0: "ElementWise", 1: "Reduce", 2: "Matmul", 3: "Unfused", 4: "ReduceBroadcasted", 5: "View"}[b.kind] for b in blocks)
    got = out_stream.numpy()
    err = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
    for _ in range(warmup):
        stream_forward()
    st.clear_blocks()
    lib.b200_launch_count_reset()
    ms, host = timed(stream_forward, steps)
    launches = int(lib.b200_launch_count()) // steps
    cached = st.blocks()
    cs = st.cache_stats()
    res["via_stream"] = {
        "ms": round(ms, 3), "host_ms": round(host, 3), "launches": launches,
        "first_call_ms": round(first_ms, 2), "blocks_per_forward": len(blocks), "block_kinds": dict(kinds),
        "ops_per_forward": int(sum(b.n_ops for b in blocks)),
        "fused_ops_in_matmul_epilogues": int(sum(b.n_ops - 1 for b in blocks if b.kind == 2)),
        "in_place_outputs": int(sum(b.aliased for b in blocks)),
        "served_from_plan_cache": bool(cached) and all(b.from_cache for b in cached),
        "plan_cache": {"hits": int(cs.hits), "misses": int(cs.misses), "plans": int(cs.plans)},
        "max_rel_diff_vs_hand_chain": err,
    }
    # ---- the cached plan captured in a CUDA graph (what a burn-fusion + graph_capture integration replays)
    try:
        with dv.Graph.capture() as graph:
            stream_forward()
        for _ in range(warmup):
            graph.launch()
        gms, ghost = timed(graph.launch, steps)
        res["via_stream_graph"] = {"ms": round(gms, 3), "host_ms": round(ghost, 3), "kernel_nodes": int(graph.kernel_nodes)}
        check(lib.b200_device_sync())
        graph.destroy()
    except abi.B200Error as e:           # reported, not hidden
        res["via_stream_graph"] = {"error": str(e)[:200]}
    xs.drop()
    st.close()
    return {"workload": f"configs[3] TransformerEncoder forward d_model {d}, {cfg['L']} layers, {cfg['h']} heads, seq {S}, batch {B}, {mm}",
            "tokens": B * S, **res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="lm", choices=sorted(CONFIGS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mm", default="tf32", choices=["tf32", "bf16", "f32x3"])
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--bucket-mb", type=int, default=32)
    ap.add_argument("--sync", default=None, choices=["nccl", "peer", "fused"],
                    help="gradient synchronisation at world > 1 (default: $B200_GRAD_SYNC, else fused)")
    ap.add_argument("--via-stream", action="store_true",
                    help="configs[3] forward through the host fusion layer's operation stream vs the hand-sequenced forward")
    args = ap.parse_args()
    if args.via_stream:
        print(json.dumps(run_via_stream(args.steps, args.warmup, args.mm)), flush=True)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run(args.config, args.steps, args.warmup, rank, world, local_rank, args.mm, not args.eager, args.bucket_mb, sync=args.sync)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""torchrun --nproc-per-node N tools/peer_gather_check.py : the NVLink peer-memory all-gathers of csrc/peer_gather.cu
(engine.PeerGather: one kernel; engine.PeerExchange: put / collect ring) against dist.all_gather_into_tensor -- equal
and unequal row counts, eager calls and CUDA-graph replays, ranks drifting apart, overflow flags, device time per call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from mccnn_b200.engine import PeerExchange, PeerGather


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rows, cols = 32, 40
    pg = PeerGather(rows, cols, dev)
    px = PeerExchange(rows, cols, dev)
    ok = True
    for step in range(50):
        x = torch.full((rows, cols), float(rank * 1000 + step), device=dev) + torch.arange(cols, device=dev)
        want = torch.empty((world * rows, cols), device=dev)
        dist.all_gather_into_tensor(want, x)
        ok &= bool(torch.equal(pg(x), want))
        ok &= bool(torch.equal(px(x), want))
    # put inside a graph, collect outside, the ranks deliberately out of step (rank r sleeps r x 2 ms before its puts)
    xs = torch.zeros((rows, cols), device=dev)
    flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        px.put(xs, flag=flag)
        px.collect()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        px.put(xs, flag=flag)
    for step in range(40):
        xs.fill_(float(rank * 7 + step))
        flag.fill_(1 if (step == 17 and rank == world - 1) else 0)
        if step % 5 == rank % 5:
            torch.cuda._sleep(4_000_000)
        g.replay()
        out, meta = px.collect()
        want = torch.empty((world * rows, cols), device=dev)
        dist.all_gather_into_tensor(want, xs)
        ok &= bool(torch.equal(out, want))
        m = meta.cpu()
        ok &= int(m[2 * world]) == world * rows and int(m[2 * world + 1]) == (1 if step == 17 else 0)
    # unequal row counts, compacted (per-point logits of a segmentation shard)
    cap = 1000 + 731 * world + 600     # the largest shard below: 1000 + 731 * (world - 1) + 97 * 5
    pc = PeerExchange(cap, 21, dev, compact=True)
    for step in range(6):
        n = 1000 + 731 * rank + 97 * step
        x = torch.randn((n, 21), device=dev)
        pc.put(x)
        out, meta = pc.collect()
        counts = torch.zeros((world,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([n], dtype=torch.int64, device=dev))
        padded = torch.zeros((cap, 21), device=dev)
        padded[:n] = x
        allp = torch.empty((world * cap, 21), device=dev)
        dist.all_gather_into_tensor(allp, padded)
        want = torch.cat([allp[r * cap: r * cap + int(counts[r])] for r in range(world)])
        total = int(meta[2 * world])
        ok &= total == want.shape[0] and bool(torch.equal(out[:total], want))
    torch.cuda.synchronize()

    def timeit(f, n=200):
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / n

    want = torch.empty((world * rows, cols), device=dev)
    big = torch.randn((cap, 21), device=dev)
    bigw = torch.empty((world * cap, 21), device=dev)
    t_nccl = timeit(lambda: dist.all_gather_into_tensor(want, xs))
    t_peer = timeit(lambda: pg(xs))
    t_px = timeit(lambda: px(xs))
    t_big_nccl = timeit(lambda: dist.all_gather_into_tensor(bigw, big), 50)
    t_big_px = timeit(lambda: pc(big), 50)
    res = torch.tensor([float(ok), -t_nccl, -t_peer, -t_px, -t_big_nccl, -t_big_px], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print({"world": world, "all_equal": bool(res[0].item() == 1.0), "nccl_us_per_call": round(-res[1].item(), 2),
               "peer_gather_us_per_call": round(-res[2].item(), 2), "peer_exchange_us_per_call": round(-res[3].item(), 2),
               "nccl_420KB_us": round(-res[4].item(), 2), "peer_exchange_420KB_us": round(-res[5].item(), 2)}, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""torchrun --nproc-per-node N tools/peer_gather_check.py : PeerGather (NVLink peer-memory all-gather, csrc/peer_gather.cu)
against dist.all_gather_into_tensor -- results over many steps (eager and CUDA-graph replays) and device time per call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from mccnn_b200.engine import PeerGather


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rows, cols = 32, 40
    pg = PeerGather(rows, cols, dev)
    ok = True
    for step in range(50):
        x = torch.full((rows, cols), float(rank * 1000 + step), device=dev) + torch.arange(cols, device=dev)
        want = torch.empty((world * rows, cols), device=dev)
        dist.all_gather_into_tensor(want, x)
        got = pg(x)
        ok &= bool(torch.equal(got, want))
    # graph replays
    xs = torch.zeros((rows, cols), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        pg(xs)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outg = pg(xs)
    for step in range(50):
        xs.fill_(float(rank * 7 + step))
        g.replay()
        want = torch.empty((world * rows, cols), device=dev)
        dist.all_gather_into_tensor(want, xs)
        ok &= bool(torch.equal(outg, want))
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
    t_nccl = timeit(lambda: dist.all_gather_into_tensor(want, xs))
    t_peer = timeit(lambda: pg(xs))
    t_graph = timeit(lambda: g.replay())
    res = torch.tensor([float(ok), t_nccl, t_peer, t_graph], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print({"world": world, "all_equal": bool(res[0].item() == 1.0), "nccl_us_per_call": round(t_nccl, 2),
               "peer_eager_us_per_call": round(t_peer, 2), "peer_graph_us_per_replay": round(t_graph, 2)}, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Where does the host time of one pipelined end-to-end step go?  Times, on the GPU box, the pieces of
PipelinedHostForward.submit_packed / result separately (host wall clock per call, queue never drained inside a
segment) and the device-side throughput of graph replays at several depths.  Diagnostic only; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_batch  # noqa: E402
from mccnn_b200 import variables  # noqa: E402
from mccnn_b200.engine import PipelinedForward, PipelinedHostForward, pack_batch  # noqa: E402
from mccnn_b200.models import MCClassS  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, P, k = 32, 1024, 64
    store = variables.VariableStore(device=dev, seed=3)

    def fn(p, b, f):
        return MCClassS.create_network(p, b, f, 1, B, k, 40, False, 1.0, 1.0, store=store)

    host = [make_batch(B, P, "sphere", seed=i) for i in range(8)]
    resident = [tuple(torch.from_numpy(a).to(dev) for a in h) for h in host]
    blocks_dev = [pack_batch(*r) for r in resident]
    blocks_pin = [pack_batch(*(torch.from_numpy(a) for a in h), pin=True) for h in host]
    out = {}
    n = 400
    for depth in (1, 2, 4, 8):
        pipe = PipelinedForward(fn, B, depth=depth)
        pipe.calibrate(resident)
        for i in range(3 * depth):
            pipe.submit_packed(blocks_dev[i % 8])
        torch.cuda.synchronize()
        # (a) device-resident submit: host cost per call (no sync) and device throughput
        t0 = time.perf_counter()
        for i in range(n):
            pipe.submit_packed(blocks_dev[i % 8])
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        # (b) bare graph replays round-robin on the lanes' streams, no input copy, no events
        t0 = time.perf_counter()
        for i in range(n):
            lane = i % depth
            with torch.cuda.stream(pipe.streams[lane]):
                pipe.engines[lane].graph.replay()
        t_host_replay = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all_replay = time.perf_counter() - t0
        # (c) host-resident: pinned H2D + replay + D2H, results consumed
        phf = PipelinedHostForward(pipe, dev)
        for rep in range(2):
            for tk in [phf.submit_packed(blocks_pin[i % 8]) for i in range(depth)]:
                phf.result(tk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tickets = []
        t_submit = 0.0
        t_result = 0.0
        for i in range(n):
            a = time.perf_counter()
            tickets.append(phf.submit_packed(blocks_pin[i % 8]))
            b = time.perf_counter()
            t_submit += b - a
            if len(tickets) > depth - 1:
                phf.result(tickets.pop(0))
                t_result += time.perf_counter() - b
        while tickets:
            phf.result(tickets.pop(0))
        t_e2e = time.perf_counter() - t0
        out["depth%d" % depth] = {
            "submit_dev_host_us": round(t_host / n * 1e6, 1), "submit_dev_total_us": round(t_all / n * 1e6, 1),
            "replay_host_us": round(t_host_replay / n * 1e6, 1), "replay_total_us": round(t_all_replay / n * 1e6, 1),
            "e2e_us": round(t_e2e / n * 1e6, 1), "e2e_submit_us": round(t_submit / n * 1e6, 1),
            "e2e_result_us": round(t_result / n * 1e6, 1)}
        if depth == 4:
            # pieces of the host path in isolation (lane 0 stream)
            s = pipe.streams[0]
            eng = pipe.engines[0]
            res_host = torch.empty(eng._result.shape, dtype=torch.float32).pin_memory()
            with torch.cuda.stream(s):
                t0 = time.perf_counter()
                for i in range(n):
                    eng._in_block.copy_(blocks_pin[i % 8], non_blocking=True)
                t_h2d_host = time.perf_counter() - t0
                s.synchronize()
                t_h2d = time.perf_counter() - t0
                t0 = time.perf_counter()
                for i in range(n):
                    res_host.copy_(eng._result, non_blocking=True)
                t_d2h_host = time.perf_counter() - t0
                s.synchronize()
                t_d2h = time.perf_counter() - t0
                t0 = time.perf_counter()
                for i in range(n):
                    ev = torch.cuda.Event()
                    ev.record(s)
                t_ev = time.perf_counter() - t0
                t0 = time.perf_counter()
                for i in range(n):
                    eng._in_block.copy_(blocks_pin[i % 8], non_blocking=True)
                    eng.graph.replay()
                    res_host.copy_(eng._result, non_blocking=True)
                    s.synchronize()
                t_serial = time.perf_counter() - t0
            out["pieces"] = {"h2d_host_us": round(t_h2d_host / n * 1e6, 1), "h2d_total_us": round(t_h2d / n * 1e6, 1),
                             "d2h_host_us": round(t_d2h_host / n * 1e6, 1), "d2h_total_us": round(t_d2h / n * 1e6, 1),
                             "event_us": round(t_ev / n * 1e6, 1), "serial_e2e_us": round(t_serial / n * 1e6, 1),
                             "block_bytes": int(blocks_pin[0].numel()), "pinned": bool(blocks_pin[0].is_pinned())}
        del pipe
    print(json.dumps(out))


if __name__ == "__main__":
    main()

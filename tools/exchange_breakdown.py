"""Where the multi-GPU step's time goes (run under torchrun, N >= 2): the pooled config-2 step with
(a) the deal only -- no peer stores, no arrival words; (b) peer stores but no arrival words;
(c) the full exchange; (d) the full exchange with the ranks aligned before every timed step.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_breakdown.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import DrivePipeline, _lib, grid_search, plan_windows, write_back  # noqa: E402
from vehiclemodelvisualodometry_b200.scheduler import PeerGather, shard_range  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
timer = bench.Timer(dev, world)
workload = "config2_single_drive_10k_32x32_w30"
n_frames, cfg, drives, _, _ = bench.pooled_drives(workload, world, dev)
n_win = world * cfg.window_count(n_frames)
fr = shard_range(drives.n_frames, world, rank)
block = int(sys.argv[1]) if len(sys.argv) > 1 else bench.DEAL_BLOCK


class Variant(DrivePipeline):
    mirrors = True
    arrival = True
    search_only = False
    local_dummy = False      # the "peer" buffers are local memory: the stores' code path without NVLink

    def _ex(self, on):
        ex = _lib.Exchange.from_buffer_copy(self.gather.exchange)
        if not on:
            ex.n_peers = 0
        elif self.local_dummy:
            if not hasattr(self, "_dummy"):
                self._dummy = [torch.empty((n_win, 64), dtype=torch.uint8, device=dev) for _ in range(ex.n_peers)]
            for i in range(ex.n_peers):
                ex.peer_records[i] = self._dummy[i].data_ptr()
        return ex

    def _search(self):
        plan_windows(self.cfg, self.drives, into=self.plan)
        self._ex_s = self._ex(self.mirrors)
        grid_search(self.cfg, self.drives, self.plan, out=self.records, exchange=self._ex_s)

    def _write_back(self):
        if self.search_only:
            return
        self._ex_w = self._ex(self.arrival)
        write_back(self.cfg, self.drives, self.plan, self.records, blend_gps=False, out=self.trajectory,
                   frame_range=self.frame_range, exchange=self._ex_w)


def run(name, mirrors, arrival, align=False, steps=40, search_only=False, local_dummy=False):
    cls = type("V", (Variant,), {"mirrors": mirrors, "arrival": arrival, "search_only": search_only,
                                 "local_dummy": local_dummy})
    sets = [PeerGather(n_win, dev, block=block) for _ in range(2)]
    pipes = [cls(cfg, drives, blend_gps=False, gather=g, frame_range=fr) for g in sets]
    st = {"n": 0}
    tok = torch.zeros(1, device=dev)

    def step():
        p = pipes[st["n"] & 1]
        st["n"] += 1
        p.run()

    if not align:
        ms = timer(step, steps, 5) / steps
    else:
        for _ in range(5):
            step()
        timer.sync()
        evs = []
        for _ in range(steps):
            timer.flush.zero_()
            dist.all_reduce(tok)                 # ranks leave the flush together
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            evs.append((a, b))
        timer.sync()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
    if rank == 0:
        print(f"{name:60s} {ms * 1e3:8.1f} us per step", flush=True)
    del pipes
    for g in sets:
        g.close()


if rank == 0:
    print(f"world {world}, {n_win} windows pooled, block {block}", flush=True)
run("search of this rank's deal only (no peer stores, no write-back)", False, False, search_only=True)
run("search of this rank's deal + the same stores into LOCAL memory", True, False, search_only=True,
    local_dummy=True)
run("search of this rank's deal + peer stores (no write-back)", True, False, search_only=True)
run("deal + peer stores + arrival words (the shipped step)", True, True)
run("the shipped step, ranks aligned before each timed step", True, True, align=True)
# one GPU's share alone, for scale: every rank searches its deal of the pool without any peer
dist.barrier()
dist.destroy_process_group()

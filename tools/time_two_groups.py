"""Would independent conformer groups on separate streams hide each other's tails?  G engines x (C/G) conformers
enqueued back to back on G streams vs one engine with C conformers."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s, sysm, n_conf, desc = bench.build_workload("cfg3")
for G in (1, 2, 4):
    engs = [GrowthEngine(0).set_system(sysm).set_backbone(s.coords) for _ in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    def run(seed):
        outs = []
        for g, (e, st) in enumerate(zip(engs, streams)):
            outs.append(e.grow(C // G, bench.BETA, seed=seed, conf_id0=g * (C // G), stream=st))
        return outs
    run(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    outs = run(1)
    for st in streams: torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    print("groups %d x %d conformers: %.1f ms, ok %.3f" % (G, C // G, e0.elapsed_time(e1), float(torch.cat([o["ok"] for o in outs]).float().mean())), flush=True)
    del engs

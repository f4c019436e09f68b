"""cProfile of MAPDeconvolver.run (cfg2, host numpy inputs): where does the fixed per-run cost go?"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import jolideco_b200 as J
from jolideco_b200 import synthetic

class A: marginalize = False; backend = None; no_graph = False; collective = "nccl"; datasets = None
wl = synthetic.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2", seed=0)
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if os.environ.get("PINNED", "0") == "1":  # host arrays in page-locked memory, as bench.py's e2e leg hands them over
    import numpy as np
    wl["datasets"] = {n: {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) else v
                          for k, v in d.items()} for n, d in wl["datasets"].items()}
mode = "joint" if wl["name"] in bench.JOINT_WORKLOADS else "sequential"
deco, comps = bench.build_run(J, wl, A, "cuda:0", n_epochs=2, mode=mode)
deco.run(datasets=wl["datasets"], components=comps)
for rep in range(int(os.environ.get('REPS', 2))):
    deco, comps = bench.build_run(J, wl, A, "cuda:0", n_epochs=epochs, mode=mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = deco.run(datasets=wl["datasets"], components=comps)
    flux = res.flux_upsampled_total
    torch.cuda.synchronize()
    print(f"run {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms for {epochs} epochs")
deco, comps = bench.build_run(J, wl, A, "cuda:0", n_epochs=epochs, mode=mode)
pr = cProfile.Profile()
pr.enable()
res = deco.run(datasets=wl["datasets"], components=comps)
flux = res.flux_upsampled_total
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

"""A/B of one layer fwd+bwd (CUDA-graph replay, same regime as bench `value`) between the current library and an older
build of it on the SAME box:  python tools/ab_layer.py C5 [C2 ...]   (child processes; DENO_LIB_PATH selects the build)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import torch
    import bench
    wl = {"C1": "burgers1d", "C2": "darcy2d", "C3": "ns2d_w64", "C4b": "pipe2d_w128", "C5": "turb3d"}[name]
    r = bench.layer_roofline(wl, torch.device("cuda", 0), 6539.9, reps=20, attribute=True)
    print(json.dumps({"us": r["layer_fwd_bwd"]["us"], "kernels": r["kernels_us"]}))


if __name__ == "__main__":
    if sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    prev = os.path.join(ROOT, "tools", "_build", "libdeno_prev.so")
    for name in sys.argv[1:]:
        for rep in range(2):
            for label, env in (("prev", {"DENO_LIB_PATH": prev, "DENO_ANALYSIS_TMA": "0", "DENO_WGRAD_TMA": "0"}), ("now", {})):
                e = dict(os.environ); e.update(env)
                r = subprocess.run([sys.executable, __file__, "child", name], capture_output=True, text=True, env=e, timeout=600)
                line = (r.stdout.strip().splitlines() or ["{}"])[-1]
                try:
                    d = json.loads(line)
                    print(f"{name} {label}: layer {d['us']:.1f} us  {d['kernels']}", flush=True)
                except Exception:
                    print(f"{name} {label}: failed rc={r.returncode} {r.stderr[-300:]}", flush=True)

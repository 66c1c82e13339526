"""A/B of one layer fwd+bwd (CUDA-graph replay, same regime as bench `value`) and of the whole training step between
environment variants of the SAME library on the SAME box:
    python tools/ab_env.py "DENO_BWD_OVERLAP=0" "DENO_BWD_OVERLAP=1" -- C2 C3 C1 C5 C4b
Each (config, variant) runs in its own child process, twice, interleaved."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WL = {"C1": "burgers1d", "C2": "darcy2d", "C3": "ns2d_w64", "C3r": "ns2d_rollout", "C4b": "pipe2d_w128", "C5": "turb3d"}


def child(name):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    wl = WL[name]
    r = bench.layer_roofline(wl, dev, 6539.9, reps=20, attribute=False)
    step, host, nparams, per_step, batch = bench.make_step(wl, dev, 0)
    ms = bench.timed_steps(step, 50, 5, 1, dev)
    bench.drop_step(step)
    print(json.dumps({"layer_us": r["layer_fwd_bwd"]["us"], "step_ms": ms / 50, "samples_s": batch * 50 / (ms * 1e-3),
                      "launches": per_step}))


if __name__ == "__main__":
    if sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    cut = sys.argv.index("--")
    variants, names = sys.argv[1:cut], sys.argv[cut + 1:]
    for name in names:
        for rep in range(2):
            for var in variants:
                e = dict(os.environ)
                for kv in var.split(","):
                    if "=" in kv:
                        k, v = kv.split("=", 1)
                        e[k] = v
                r = subprocess.run([sys.executable, __file__, "child", name], capture_output=True, text=True, env=e, timeout=900)
                line = (r.stdout.strip().splitlines() or ["{}"])[-1]
                try:
                    d = json.loads(line)
                    print(f"{name} [{var}]: layer {d['layer_us']:.1f} us, step {d['step_ms']:.3f} ms, "
                          f"{d['samples_s']:.0f} samples/s, {d['launches']} launches/step", flush=True)
                except Exception:
                    print(f"{name} [{var}]: failed rc={r.returncode} {r.stderr[-400:]}", flush=True)

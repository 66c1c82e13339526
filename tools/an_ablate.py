"""Timing-only ablations of plane_analysis_tc (DENO_AN_ABLATE bits; results are wrong on purpose): which phase does the
kernel's duration actually depend on?  The bits are compiled in only with -DDENO_AN_ABLATE_BUILD:
    nvcc <flags of deno4pytorch_b200/build.py> -DDENO_AN_ABLATE_BUILD -o tools/_build/libdeno_ablate.so deno4pytorch_b200/csrc/capi.cu
    DENO_LIB_PATH=tools/_build/libdeno_ablate.so python tools/an_ablate.py C2 C5   (each variant in its own child process)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BITS = [(0, "baseline"), (1, "no stage-1 MMAs"), (2, "no stage-2 MMAs"), (3, "no MMAs"), (4, "no A-image stores"), (8, "no B2 stores"),
        (16, "no spectrum stores"), (32, "no gz stores"), (64, "no staged-plane loads"), (4 + 64, "no transform loads+stores"),
        (4 + 8 + 16 + 32 + 64, "workers: waits + TMEM reads only"), (127, "everything off")]


def child(name):
    import torch
    import bench
    wl = {"C2": "darcy2d", "C3": "ns2d_w64", "C5": "turb3d"}[name]
    r = bench.layer_roofline(wl, torch.device("cuda", 0), 6539.9, reps=10, attribute=True)
    k = r["kernels_us"]
    print(json.dumps({"an": k.get("plane_analysis_tc"), "gz": k.get("plane_analysis_gz_tc")}))


if __name__ == "__main__":
    if sys.argv[1] == "child":
        child(sys.argv[2]); sys.exit(0)
    for name in sys.argv[1:]:
        for bits, label in BITS:
            e = dict(os.environ); e["DENO_AN_ABLATE"] = str(bits)
            r = subprocess.run([sys.executable, __file__, "child", name], capture_output=True, text=True, env=e, timeout=600)
            line = (r.stdout.strip().splitlines() or ["{}"])[-1]
            try:
                d = json.loads(line)
                print(f"{name} ablate={bits:3d} {label:36s} analysis {d['an']:7.2f} us   analysis_gz {d['gz']:7.2f} us", flush=True)
            except Exception:
                print(f"{name} ablate={bits}: failed {r.stderr[-300:]}", flush=True)

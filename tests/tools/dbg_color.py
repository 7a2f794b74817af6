"""Developer tool (not a test): `--run` executes colour cases through the C ABI on the GPU box
and stores the outputs in gpurun_out/dbg_color.npz; `--check` (CPU, here) compares them with
the oracle and prints error statistics per case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import pipeline_oracle as P  # noqa: E402
import plbtest as T  # noqa: E402
import test_gpu_color as G  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "..", "gpurun_out", "dbg_color.npz")

HDR10, SDR = G.HDR10, G.SDR
SDR2020 = T.csp(T.PRIM["BT_2020"], T.TRC["BT_1886"])
LIN2020 = T.csp(T.PRIM["BT_2020"], T.TRC["LINEAR"])
LIN709 = T.csp(T.PRIM["BT_709"], T.TRC["LINEAR"])


def cases():
    c = {}
    for hq in (0, 1):
        c[f"cmap_default_hq{hq}"] = (HDR10, SDR, hq, None, None, G.rgb(4))
    for tone in T.TONE_FNS:
        c[f"tone_{tone}"] = (HDR10, SDR, 0, tone, "clip", G.rgb(5))
    for gm in T.GAMUT_FNS:
        c[f"gamut_{gm}"] = (SDR2020, SDR, 0, "clip", gm, G.rgb(6))
    c["lin_gamut_clip"] = (LIN2020, LIN709, 0, "clip", "clip", G.rgb(6))
    c["lin_gamut_perceptual"] = (LIN2020, LIN709, 0, "clip", "perceptual", G.rgb(6))
    c["inverse"] = (SDR, HDR10, 0, None, None, G.rgb(7))
    return c


def run():
    import gpu_helpers
    out = {}
    with gpu_helpers.Gpu() as gpu:
        for name, (s, d, hq, tone, gm, src) in cases().items():
            oa = [s, d, [float(hq), T.TONE_FNS.index(tone) if tone else -1,
                         T.GAMUT_FNS.index(gm) if gm else -1, 1.0]]
            try:
                out[name] = G.run_ops(gpu, src, [5], oa)
            except Exception as e:  # noqa: BLE001
                print(name, "FAILED", e)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("saved", len(out))


def check():
    got = np.load(OUT)
    orc, t0 = T.oracle(), T.t0()
    for name, (s, d, hq, tone, gm, src) in cases().items():
        if name not in got:
            print(f"{name:28s} missing")
            continue
        cm = P.color_map_params(t0, s.copy(), d.copy(), hq=bool(hq), tone_fn=tone, gamut_fn=gm)
        want = orc.color_map(src, cm["params"]) if cm is not None else src
        g = got[name]
        dd = np.abs(g[..., :3] - want[..., :3])
        bad = np.unravel_index(np.argmax(dd), dd.shape)
        print(f"{name:28s} max {dd.max():.3e} mean {dd.mean():.3e} frac>1e-5 {np.mean(dd > 1e-5):.4f} "
              f"worst in {src[bad[0], bad[1], :3]} got {g[bad[0], bad[1], :3]} want {want[bad[0], bad[1], :3]}")


if __name__ == "__main__":
    run() if "--run" in sys.argv else check()

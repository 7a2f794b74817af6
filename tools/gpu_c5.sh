timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_render.py -q -m gpu -x --timeout 300 -k "error_diffusion or c5" 2>&1 | tail -5
for v in 0 1; do
PLB_ERRDIFF_CLUSTER=$v timeout 300 python - <<'PY'
import os, json, sys
sys.argv=["bench.py"]
import bench, torch
class A: pass
cx = bench.Ctx(); cx.torch=torch; cx.dist=None; cx.rank=0; cx.world=1; cx.local_rank=0; cx.want_parity=False
torch.cuda.set_device(0)
for lanes in (1, 8, 16, 18):
    r = bench.measure_lanes(cx, "c5_8k_dovi_errdiff", lanes, 2)
    print("CLUSTER", os.environ.get("PLB_ERRDIFF_CLUSTER"), "lanes", lanes, r["value"], r["one_frame_at_a_time"])
PY
done

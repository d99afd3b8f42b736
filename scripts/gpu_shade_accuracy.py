"""Bring-up helper: image error of the GPU renderer against the oracle for the library selected with VKRT_LIB_PATH.
Prints per-channel relative L2 (all pixels / without outliers), outlier fraction and NaN counts on three small scenes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from vk_raytrace_b200 import Context, scenes, structs as S, upload_scene
from oracle import binding as ob
from helpers import small_sponza


def pc_rel_l2(a, b):
    return max(float(np.sqrt(((a[..., c] - b[..., c]) ** 2).sum() / max((b[..., c] ** 2).sum(), 1e-30))) for c in range(3))


def main():
    g = Context(0)
    cases = [("cornell d4", scenes.cornell_box(with_light=True), 64, 64, 4, 4, {}),
             ("sponza d6 pbr0", small_sponza(), 96, 64, 6, 4, {}),
             ("sponza d6 pbr1", small_sponza(), 96, 64, 6, 4, dict(pbrMode=1)),
             ("helmet sunsky d4", scenes.helmet_standin(level=3, tex_size=64, env_size=(64, 32)), 96, 64, 4, 4, dict(sunsky=True))]
    for name, sc, W, H, depth, frames, kw in cases:
        if sc is None:
            continue
        o = ob.Oracle()
        upload_scene(g, sc, W, H); upload_scene(o, sc, W, H)
        if kw.get("sunsky"):
            ss = S.default_sunsky(in_use=1)
            g.set_sunsky(ss); o.set_sunsky(ss)
        st = S.default_rtx_state(W, H, maxDepth=depth, pbrMode=kw.get("pbrMode", 0))
        g.render(st, frames); o.render(st, frames)
        a, b = g.output_read()[..., :3], o.output_read()[..., :3]
        off = (np.abs(a - b) > 1e-3 * (1 + np.abs(b))).any(axis=2)
        print(f"{name:18s} nan={int((~np.isfinite(a)).sum())} relL2_all={pc_rel_l2(a, b):.2e} off={off.mean():.4f} relL2_rest={pc_rel_l2(a[~off], b[~off]):.2e} "
              f"maxabs_rest={np.abs(a[~off] - b[~off]).max():.2e}", flush=True)
        for dbg in (1, 5, 9, 10, 11):
            st2 = S.default_rtx_state(W, H, maxDepth=depth, debugging_mode=dbg)
            g.output_clear()
            g.render(st2, 1); o.render(st2, 1)
            a, b = g.output_read()[..., :3], o.output_read()[..., :3]
            print(f"   debug {dbg:2d}: maxabs={np.nanmax(np.abs(a - b)):.2e} relL2={pc_rel_l2(a, b):.2e} nan={int((~np.isfinite(a)).sum())}", flush=True)


main()

"""Generate tests/golden/ref_golden.npz by EXECUTING THE UNMODIFIED REFERENCE (oracle/_ref, i.e. the
headers under /root/reference compiled against oracle/shim). Run in the build container only:

    python tests/golden/make_golden.py

The reference's own tests pin no pixel of imprint / texture deposit / whole-image compose
(SURVEY.md §4), so these fixtures are the pinned authority that travels to the GPU box.
Contents: inputs (seeded), checksums and crops of the outputs for
  gui_*   : painty_gui style 3-point footprint stroke, r=30, 768x1024
  tex_*   : TextureBrushTest stroke (r=40) + a crossing stroke (r=25), smudge off
  texs_*  : TextureBrushTest exactly as the reference runs it (smudge on, the CPU class's default)
  km_*    : 4096 random pixels through Renderer::compose incl. edge cases (d=0, S=0, K=0 -> NaN)
  brd_*   : footprint stroke overhanging the top-left border with fractional centres (B#11), r=11 (an OOB-free radius)
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle.cpu import Cpu  # noqa: E402
from tests.workloads import gui_stroke_imprints, km_random_planes  # noqa: E402


def main():
    ref = Cpu("ref")
    out = {}
    # --- gui stroke
    K, S = ref.compute_scattering_absorption([.2, .05, .4], [.6, .3, .7])
    cx, cy, th = gui_stroke_imprints(ref, [(100.3, 200.7), (400.9, 260.2), (700.1, 180.4)])
    cv = ref.canvas(768, 1024)
    br = ref.footprint_brush(30.0)
    br.dip(K, S)
    br.imprint_batch(cv, cx, cy, th)
    R = cv.compose()
    st = cv.get()
    pK, pS, pV = br.pickup_map()
    out.update(gui_paint_K=K, gui_paint_S=S, gui_cx=cx, gui_cy=cy, gui_theta=th, gui_sumR=R.sum(), gui_sumV=st["V"].sum(),
               gui_wet=(st["V"] > 0).sum(), gui_R_230_400=R[230, 400], gui_R_crop=R[200:264, 380:444].copy(),
               gui_V_crop=st["V"][200:264, 380:444].copy(), gui_K_crop=st["K"][200:264, 380:444].copy(),
               gui_pickV=pV, gui_pickK=pK, gui_R_rowsum=R.sum(axis=(1, 2)), gui_R_colsum=R.sum(axis=(0, 2)))
    # --- texture strokes
    cv = ref.canvas(768, 1024)
    tb = ref.texture_brush()
    tb.set_radius(40.0)
    tb.dip([.2, .3, .4], [.1, .23, .14])
    tb.paint_stroke(cv, [(50, 250), (400, 250), (650, 250)])
    tb.dip([.5, .1, .2], [.3, .2, .5])
    tb.set_radius(25.0)
    tb.paint_stroke(cv, [(300.5, 50.2), (350.1, 200.7), (330.3, 400.9), (420.0, 600.5)])
    R = cv.compose()
    st = cv.get()
    out.update(tex_sumR=R.sum(), tex_sumV=st["V"].sum(), tex_wet=(st["V"] > 0).sum(), tex_R_crop=R[200:300, 300:400].copy(),
               tex_V_crop=st["V"][200:300, 300:400].copy(), tex_R_rowsum=R.sum(axis=(1, 2)), tex_R_colsum=R.sum(axis=(0, 2)))
    # --- the reference's TextureBrushTest as written: smudge ON (the CPU class's default), radius 40
    cv = ref.canvas(768, 1024)
    tb = ref.texture_brush()
    tb.enable_smudge(True)
    tb.dip([.2, .3, .4], [.1, .23, .14])
    tb.set_radius(40.0)
    tb.paint_stroke(cv, [(50, 250), (400, 250), (650, 250)])
    R = cv.compose()
    st = cv.get()
    out.update(texs_sumR=R.sum(), texs_sumV=st["V"].sum(), texs_wet=(st["V"] > 0).sum(), texs_R_crop=R[200:300, 300:400].copy(),
               texs_V_crop=st["V"][200:300, 300:400].copy(), texs_R_colsum=R.sum(axis=(0, 2)))
    # --- compose, random + edge cases
    Kp, Sp, Vp, R0p = km_random_planes(64, 64, seed=7, edge_cases=True)
    out.update(km_K=Kp, km_S=Sp, km_V=Vp, km_R0=R0p, km_R=ref.compose(Kp, Sp, Vp, R0p))
    # --- border overhang, fractional centres, snapshot on, two dips
    cv = ref.canvas(96, 128)
    br = ref.footprint_brush(11.0)
    br.dip([.3, .2, .1], [.2, .4, .3])
    n = 40
    bx = np.linspace(3.4, 40.2, n)
    by = np.linspace(2.6, 30.9, n)
    bt = np.linspace(-2.0, 2.5, n)
    br.imprint_batch(cv, bx, by, bt)
    br.dip([.1, .5, .2], [.4, .1, .3])
    br.imprint_batch(cv, bx[::-1].copy(), by[::-1].copy(), bt)
    st = cv.get()
    out.update(brd_cx=bx, brd_cy=by, brd_theta=bt, brd_K=st["K"], brd_S=st["S"], brd_V=st["V"], brd_R=cv.compose())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("gui sumR", repr(out["gui_sumR"]), "sumV", repr(out["gui_sumV"]), "wet", out["gui_wet"])


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY. Bake the reference's *input* assets into one small fixture.

Run in the build container (where /root/reference exists):  python oracle/bake_assets.py
Writes tests/golden/assets/painty_assets.npz with the raw integer pixel data of
  * data/footprint/footprint.png      (1024x1024 gray u8; FootprintBrush.hxx:49)
  * data/sample_0/thickness_map.png   (171x800 gray u16; BrushStrokeSample.cxx:164)
  * data/textures/*.png               (236 brush textures, gray u16 whose two bytes are equal -> stored as u8;
                                       renderer/src/TextureBrushDictionary.cxx:81-118) -> painty_textures.npz
  * data/canvas_patterns/0.png        (2048x2048 RGB u8; renderer/src/CanvasGpu.cxx:27-40) -> painty_textures.npz
and the three JSON palettes (mixer/src/Serialization.cxx:26-53 format [{"K":[3],"S":[3]}]) as
f64 arrays. No reference *source* is copied; these are the data files the hot path consumes.
Linearisation (Color.hxx:189-195) and LANCZOS4 resizing are done at load time by
painty_b200/assets.py with the container's cv2, so oracle and device consume identical bytes.
"""
import json
import os

import cv2
import numpy as np

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "assets")


def main():
    os.makedirs(OUT, exist_ok=True)
    fp = cv2.imread(f"{REF}/footprint/footprint.png", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_GRAYSCALE)
    tm = cv2.imread(f"{REF}/sample_0/thickness_map.png", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_GRAYSCALE)
    assert fp.dtype == np.uint8 and fp.shape == (1024, 1024)
    assert tm.dtype == np.uint16 and tm.shape == (171, 800)
    pal = {}
    for name in ("curtis-watercolor", "lindemeier-measured", "thinning_medium"):
        j = json.load(open(f"{REF}/paint_palettes/{name}.json"))
        pal[name.replace("-", "_") + "_K"] = np.array([p["K"] for p in j], dtype=np.float64)
        pal[name.replace("-", "_") + "_S"] = np.array([p["S"] for p in j], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "painty_assets.npz"), footprint_u8=fp, thickness_u16=tm, **pal)
    print("wrote", os.path.join(OUT, "painty_assets.npz"), os.path.getsize(os.path.join(OUT, "painty_assets.npz")), "bytes")
    tex = {}
    for name in sorted(os.listdir(f"{REF}/textures")):
        g = cv2.imread(f"{REF}/textures/{name}", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_GRAYSCALE)
        assert g.dtype == np.uint16 and ((g >> 8) == (g & 0xFF)).all()  # v * 257: the u8 high byte restores the u16 exactly
        tex["tex_" + name[:-4]] = (g >> 8).astype(np.uint8)
    pat = cv2.cvtColor(cv2.imread(f"{REF}/canvas_patterns/0.png", cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    assert pat.dtype == np.uint8 and pat.shape == (2048, 2048, 3)
    path = os.path.join(OUT, "painty_textures.npz")
    np.savez_compressed(path, canvas_pattern_u8=pat, **tex)
    print("wrote", path, os.path.getsize(path), "bytes,", len(tex), "textures")


if __name__ == "__main__":
    main()

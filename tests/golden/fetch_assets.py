"""Copies the two small image ASSETS (data, not source) the reference's samples load into tests/golden/assets/, so that the
GPU box - which has no /root/reference - renders the same texels:

  resources/font/font_enu.png                    400x400 RGB   samples/AnisotropicFilter/AnisotropicFilter.cpp:203
  resources/texture_and_blending/chessboard.png  32x32 RGBA    samples/TextureAndBlending/TextureAndBlending.cpp:226

(Dirt.jpg and the Sponza textures are Git-LFS pointers upstream; seeded procedural textures stand in for them.)
Run here, where /root/reference exists:  python tests/golden/fetch_assets.py
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/resources"
FILES = {"font_enu.png": "font/font_enu.png", "chessboard.png": "texture_and_blending/chessboard.png"}

if __name__ == "__main__":
    out = os.path.join(HERE, "assets")
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name, rel in FILES.items():
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(out, name))
        os.chmod(os.path.join(out, name), 0o644)
        manifest[name] = {"source": "resources/" + rel,
                          "sha256": hashlib.sha256(open(os.path.join(out, name), "rb").read()).hexdigest()}
    json.dump(manifest, open(os.path.join(out, "MANIFEST.json"), "w"), indent=1)
    print(manifest)

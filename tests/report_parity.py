"""Developer tool: renders the parity scenes on two backends and prints per-buffer mismatch counts.

    python tests/report_parity.py [--a cuda|oracle|ref] [--b ref|oracle] [--only soup,c1,c2,probe]
"""
import argparse
import faulthandler
import itertools
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from salviarenderer_b200 import abi as A, scenes as S  # noqa: E402

LIBS = {
    "cuda": os.path.join(ROOT, "salviarenderer_b200", "csrc", "libsalvia_b200.so"),
    "oracle": os.path.join(ROOT, "oracle", "libsalvia_oracle.so"),
    "ref": os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so"),
}
BAD = []


def cmp(ra, rb, tag, color_tol=0):
    ok = True
    msgs = []
    for k in ("color", "depth", "stencil", "resolved", "count"):
        a, b = getattr(ra, k), getattr(rb, k)
        if a is None:
            continue
        if k == "depth":
            a, b = a.view(np.uint32), b.view(np.uint32)
        if k in ("color", "resolved") and color_tol:
            diff = np.abs(a.astype(np.int32) - b.astype(np.int32))
            nd = int((diff > color_tol).sum())
            n1 = int((diff > 0).any(axis=-1).sum())
            if n1:
                msgs.append(f"{k}: {n1} px differ by <= {int(diff.max())} LSB")
        else:
            nd = int((a != b).sum())
        if nd:
            ok = False
            idx = np.argwhere(a != b)[:3]
            msgs.append(f"{k} DIFF {nd} first {idx.tolist()} a={a[tuple(idx[0])]} b={b[tuple(idx[0])]}")
    sa = {k: v for k, v in ra.stats.items() if k not in ("vs_invocations", "gs_invocations")}
    sb = {k: v for k, v in rb.stats.items() if k not in ("vs_invocations", "gs_invocations")}
    if sa != sb:
        ok = False
        msgs.append(f"stats {sa} vs {sb}")
    print(("OK   " if ok else "FAIL ") + tag + ("  | " + " ; ".join(msgs) if msgs else ""), flush=True)
    if not ok:
        BAD.append(tag)
    return ok


def both(bea, beb, mk, tag, frames=(0,), color_tol=0):
    a, b = mk(), mk()
    a.setup(bea)
    b.setup(beb)
    for f in frames:
        t = time.time()
        ra = a.run(bea, f)
        t1 = time.time()
        rb = b.run(beb, f)
        t2 = time.time()
        cmp(ra, rb, f"{tag} f{f} ({t1 - t:.3f}s/{t2 - t1:.3f}s) cprims={rb.stats['cprimitives']} ps={rb.stats['ps_invocations']}", color_tol)


def probes(bea, beb):
    rng = np.random.default_rng(5)
    n = 4000
    coords = rng.uniform(-2.5, 3.5, size=(n, 2)).astype(np.float32)
    coords[:50] = rng.integers(-3, 4, size=(50, 2)).astype(np.float32) * 0.5
    scale = (10.0 ** rng.uniform(-4, -0.5, size=(n, 1))).astype(np.float32)
    ddx = (rng.normal(size=(n, 2)) * scale).astype(np.float32)
    ddy = (rng.normal(size=(n, 2)) * scale * rng.uniform(0.05, 1, size=(n, 1))).astype(np.float32)
    ddx[:10] = 0
    ddy[:10] = 0
    lod = rng.uniform(-2, 11, size=n).astype(np.float32)
    imgs = {"noise64x32": S.noise_texture(64, 3)[:32], "chess32": S.chessboard_texture()}
    bad = total = 0
    for name, img in imgs.items():
        for fmt in (A.PF_RGBA8, A.PF_RGBA32F, A.PF_RG32F, A.PF_BGRA8):
            if fmt == A.PF_RGBA32F:
                data = (img.astype(np.float32) / 255).astype(np.float32)
            elif fmt == A.PF_RG32F:
                data = np.ascontiguousarray(img[..., :2].astype(np.float32) / 255)
            else:
                data = img
            ta, tb = S.make_texture(bea, data, fmt), S.make_texture(beb, data, fmt)
            for l in range(bea.level_count(ta)):
                if not np.array_equal(bea.read_texture(ta, l), beb.read_texture(tb, l)):
                    print("MIP DIFF", name, fmt, l)
                    bad += 1
            for minf, magf, mipf in itertools.product((0, 1), (0, 1), (0, 1, 2)):
                for au, av in ((0, 0), (1, 1), (2, 2), (3, 3), (0, 2), (2, 1), (1, 0), (3, 0)):
                    if 3 in (au, av) and (minf == 1 or magf == 1):
                        continue
                    if fmt == A.PF_BGRA8 and (minf == 1 or magf == 1):
                        continue
                    for q in (0, 1, 2):
                        for an in ((16, 4) if mipf == 2 else (0,)):
                            d = A.sampler_desc(minf, magf, mipf, q, au, av, an, border=(0.25, 0.5, 0.75, 1.0))
                            sa, sb = bea.create_sampler(d, ta), beb.create_sampler(d, tb)
                            a = bea.sampler_probe(sa, coords, ddx, ddy)
                            b = beb.sampler_probe(sb, coords, ddx, ddy)
                            total += 1
                            if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                                i = np.argwhere((a.view(np.uint32) != b.view(np.uint32)).any(1))[:, 0]
                                bad += 1
                                print("GRAD DIFF", name, fmt, minf, magf, mipf, au, av, q, an, len(i), a[i[0]], b[i[0]], coords[i[0]], ddx[i[0]], ddy[i[0]])
                            if mipf != 2:
                                a = bea.sampler_probe(sa, coords, lod=lod)
                                b = beb.sampler_probe(sb, coords, lod=lod)
                                if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                                    i = np.argwhere((a.view(np.uint32) != b.view(np.uint32)).any(1))[:, 0]
                                    bad += 1
                                    print("LOD DIFF", name, fmt, minf, magf, mipf, au, av, q, len(i), a[i[0]], b[i[0]], coords[i[0]], lod[i[0]])
    print(("OK   " if not bad else "FAIL ") + f"sampler probes: {total} configs, {bad} bad", flush=True)
    if bad:
        BAD.append("probes")


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", default="cuda")
    ap.add_argument("--b", default="ref")
    ap.add_argument("--only", default="c1,soup,c2,probe")
    args = ap.parse_args()
    bea, beb = A.Backend(LIBS[args.a]), A.Backend(LIBS[args.b])
    print("backends:", bea.name, "vs", beb.name, flush=True)
    which = args.only.split(",")
    B = lambda mk, tag, frames=(0,), tol=0: both(bea, beb, mk, tag, frames, tol)  # noqa: E731
    if "c1" in which:
        for s in (1, 4):
            B(lambda: S.ColorizedTriangle(800, 600, s, with_count=True), f"C1 s{s}", range(2))
    if "soup" in which:
        for s in (1, 2, 4):
            for cull in (A.CULL_NONE, A.CULL_BACK, A.CULL_FRONT):
                B(lambda: S.TriangleSoup(samples=s, cull=cull, seed=7 + s), f"soup s{s} cull{cull}")
        B(lambda: S.TriangleSoup(samples=4, strip=True, n=500), "strip")
        B(lambda: S.TriangleSoup(samples=1, index_dtype=np.uint32, n=2000, size=0.2, w=512, h=512), "small tris")
        B(lambda: S.TriangleSoup(samples=1, n=3000, size=0.1, w=1000, h=600), "odd target size")
        B(lambda: S.TriangleSoup(samples=4, modifiers=[A.AM_CENTROID | A.AM_LINEAR]), "centroid")
        B(lambda: S.TriangleSoup(samples=4, modifiers=[A.AM_NOPERSPECTIVE]), "noperspective")
        B(lambda: S.TriangleSoup(samples=2, modifiers=[A.AM_NOINTERPOLATION]), "nointerp")
        for fn in range(8):
            B(lambda: S.TriangleSoup(samples=4, ds=A.depth_stencil_desc(depth_func=fn)), f"depthfunc{fn}")
        B(lambda: S.TriangleSoup(samples=4, ds=A.depth_stencil_desc(depth_enable=False)), "nodepth")
        B(lambda: S.TriangleSoup(samples=4, ds=A.depth_stencil_desc(depth_write=False)), "nodepthwrite")
        B(lambda: S.TriangleSoup(samples=4, bs=A.BS_LERP_SRC_ALPHA), "blend")
        B(lambda: S.TriangleSoup(samples=1, bs=A.BS_LERP_SRC_ALPHA, color_fmt=A.PF_BGRA8), "blend bgra")
        B(lambda: S.TriangleSoup(samples=1, bs=A.BS_LERP_SRC_ALPHA, color_fmt=A.PF_RGBA32F), "blend f32")
        B(lambda: S.TriangleSoup(samples=4, ps=A.PS_DISCARD_ALL), "discard")
        for sop in range(1, 9):
            for fn in (A.CMP_ALWAYS, A.CMP_LESS_EQUAL, A.CMP_NOT_EQUAL):
                ds = A.depth_stencil_desc(stencil_enable=True, read_mask=0x0F, write_mask=0x3F,
                                          front=(A.SOP_KEEP, A.SOP_KEEP, sop, fn),
                                          back=(A.SOP_ZERO, A.SOP_INVERT, (sop % 8) + 1, A.CMP_GREATER_EQUAL))
                B(lambda: S.TriangleSoup(samples=2, ds=ds, stencil_ref=5), f"stencil op{sop} fn{fn}")
    if "c2" in which:
        B(lambda: S.TextureAndBlending(640, 360), "C2 640x360", range(5))
        B(lambda: S.TextureAndBlending(1920, 1080), "C2 1080p", (0, 3))
        B(lambda: S.TextureAndBlending(640, 360, samples=4, ps_program=A.PS_TEX_GRAD_ALPHA, mip_filter=A.FILTER_ANISOTROPIC,
                                       max_aniso=16), "C2 aniso16 4x", (0, 2))
    if "probe" in which:
        probes(bea, beb)
    print("SUMMARY:", "ALL OK" if not BAD else f"{len(BAD)} FAILED: {BAD[:20]}")
    return 1 if BAD else 0


if __name__ == "__main__":
    sys.exit(main())

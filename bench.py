#!/usr/bin/env python
"""bench.py — frames/s (and shaded Mpix/s) of the north_star target: the Sponza-class scene at 3840x2160, 4x MSAA + resolve,
SASL vertex + pixel shaders compiled at run time, 16x anisotropic samplers (BASELINE.json configs[3]) on N B200s, through the
C ABI of the CUDA product.

    python bench.py --gpus N --steps K --warmup W            # the product (torchrun launches N ranks for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU renderer (rank 0 only)

A "step" is one frame: clear colour + clear depth/stencil + 24 draws (one per material group) + MSAA resolve (+ for N>1 the
sort-first assembly of the owned tiles on rank 0 over NVLink).  All inputs are resident in HBM when the timed region starts
(`value`); `e2e` repeats the measurement with the frame's geometry uploaded from pinned host memory and the resolved frame
read back to pinned host memory inside the timed region, every step.

The reference's SASL compiler (Boost.Wave / Spirit + LLVM MCJIT) cannot be built here, so the reference arm runs the cpp twins
of the two shaders (SLV_VS_SPONZA / SLV_PS_SPONZA_GRAD: sample_2d_grad with the SASL per-row / per-column derivatives) on the
unmodified reference core (oracle/_ref); `parity` in the product's line is the buffer-by-buffer comparison of one full-size
frame of the product (SASL shaders) with that reference (untimed).  `--shaders builtin --aniso 0` is round 1's headline
(samples/Sponza's cpp shaders, trilinear), reported under `variants` by default.

Timing: CUDA events recorded on the stream the kernels run on, barrier + synchronize on both sides, max over ranks.  Inputs are
larger than L2 (the 4K 4xMSAA colour + depth/stencil targets alone are 398 MB vs 126 MB of L2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GATED = ("ia_vertices", "ia_primitives", "cinvocations", "cprimitives", "ps_invocations", "backend_input_pixels")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--tex-size", type=int, default=1024)
    ap.add_argument("--aniso", type=int, default=16, help="max anisotropy (0 = trilinear, as samples/Sponza)")
    ap.add_argument("--shaders", default="sasl", choices=["sasl", "builtin"],
                    help="sasl: SASL vertex + pixel shader compiled at run time (tex2D = sample_2d_grad); builtin: samples/Sponza's cpp shaders")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size comparison with the reference renderer")
    ap.add_argument("--no-variants", action="store_true", help="skip the other shader / filter variants of the workload")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs (N = 1 only)")
    ap.add_argument("--cpu-baseline-frames", type=int, default=2)
    ap.add_argument("--parity-frame", type=int, default=3)
    ap.add_argument("--dump-frame", default=None, help="(reference arm) write the buffers of --parity-frame to this .npz")
    ap.add_argument("--scene", default="sponza", help="(reference arm) sponza | c1 | c2 | c3a | c3b | c5")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max((float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- the workload -----------------------------------------------------------------------------------------------------------
SASL_VS_SPONZA = """
float4x4 wvpMatrix; float4 lightPos; float4 eyePos;
struct VSIn  { float4 pos: POSITION; float4 tex: TEXCOORD0; float4 norm: NORMAL; };
struct VSOut { float4 pos: sv_position; float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };
VSOut vs_main(VSIn in) {
    VSOut o;
    o.norm = in.norm; o.pos = mul(in.pos, wvpMatrix); o.lightDir = lightPos - in.pos; o.eyeDir = eyePos - in.pos; o.tex = in.tex;
    return o;
}
"""
SASL_PS_SPONZA = """
sampler texSamp;
struct PSIn { float4 tex: TEXCOORD0; float4 norm: TEXCOORD1; float4 lightDir: TEXCOORD2; float4 eyeDir: TEXCOORD3; };
float4 ps_main(PSIn in): COLOR {
    float4 diff = tex2D(texSamp, in.tex.xy);
    float illum = clamp(dot(normalize(in.lightDir.xyz), normalize(in.norm.xyz)), 0.0f, 1.0f);
    return float4(diff.xyz * illum, 1.0f);
}
"""


def workload_name(args):
    sh = "SASL vertex + pixel shaders (tex2D = sample_2d_grad)" if args.shaders == "sasl" else "samples/Sponza's cpp shaders (tex2d: LOD once per quad)"
    flt = f"{args.aniso}x anisotropic" if args.aniso > 1 else "trilinear"
    return (f"Sponza-like atrium (262,249 tris, 24 material draws), {args.width}x{args.height}, {args.samples}x MSAA + resolve, "
            f"{sh}, {flt}")


def config_dict(args, n):
    """Identical for the product arm and the reference arm of the same command line (the driver compares them)."""
    return {"workload": workload_name(args),
            "width": args.width, "height": args.height, "msaa": args.samples, "triangles": 262249, "draws_per_frame": 24,
            "shaders": args.shaders, "max_anisotropy": args.aniso,
            "texture": f"24 x {args.tex_size}^2 rgba8 + mips, wrap, " + (f"{args.aniso}x anisotropic" if args.aniso > 1 else "trilinear"),
            "color_format": "bgra8", "depth_stencil_format": "rg32f",
            "parallelism": "single GPU" if n == 1 else f"sort-first: 64x64 screen tiles interleaved over {n} GPUs, geometry replicated, "
                           "finished tiles resolved straight into rank 0's surface over NVLink peer memory (fallback: NCCL gather)",
            "l2_policy": "inputs larger than L2 (398 MB of render targets per frame vs 126 MB L2); no explicit flush"}


def twin_scene(args, scenes, A):
    """The workload with built-in device programs / the reference's cpp shaders: for --shaders sasl the twins of the SASL pair."""
    if args.shaders == "sasl":
        return scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=args.aniso,
                                 ps_program=A.PS_SPONZA_GRAD, sasl_derivatives=True)
    return scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=args.aniso)


def install_sasl_shaders(sc, be, A, vs=True, ps=True):
    """Compiles the SASL pair at run time (salviarenderer_b200/sasl) and binds it to every draw of the scene."""
    from salviarenderer_b200.sasl import jit
    if vs:
        vsh = jit.compile(SASL_VS_SPONZA, "vs")
        vs_mod = jit.load(be, vsh)
        sc.vs_binding = lambda wvp, light, eye: A.shader_binding(A.program_jit(vs_mod), vsh.unit.pack_uniforms(
            {"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "lightPos": light, "eyePos": eye}))
    if ps:
        psh = jit.compile(SASL_PS_SPONZA, "ps")
        ps_mod = jit.load(be, psh)
        base = sc.frame_draws.__func__ if hasattr(sc.frame_draws, "__func__") else None

        def frame_draws(be_, frame, _cache={}):
            key = (id(be_), frame)
            if key not in _cache:
                ds = base(sc, be_, frame)
                for d, (m, _, _) in zip(ds, sc.groups):
                    d.ps = A.shader_binding(A.program_jit(ps_mod), b"", [sc.samplers[m]])
                _cache[key] = ds
            return _cache[key]

        sc.frame_draws = frame_draws
    sc._draw_cache = {}


def small_scene(name, scenes):
    """The other BASELINE.json configs at their full sizes (parity-test cases; reported under `configs`)."""
    if name == "c1":
        return scenes.ColorizedTriangle(800, 600, 1), "configs[0] ColorizedTriangle 800x600, no MSAA, cpp shaders"
    if name == "c2":
        return scenes.TextureAndBlending(1920, 1080), "configs[1] TextureAndBlending 1920x1080: trilinear + alpha blending + depth test"
    if name == "c3a":
        return scenes.ColorizedTriangle(1920, 1080, 4), "configs[2] AntiAliasing 1920x1080, 4x MSAA + resolve"
    if name == "c3b":
        return scenes.AnisotropicFilter(1920, 1080, 4), "configs[2] AnisotropicFilter 1920x1080, 4x MSAA + resolve, font_enu.png 400x400, 7 filter rows (AF <= 16x)"
    if name == "c5":
        return (scenes.HeightFieldTwoPass(7680, 4320, 1, nx=2500, nz=2000, shadowed=True),
                "configs[4] StandardShadowMap two-pass over the synthetic 10,000,000-triangle mesh at 7680x4320")
    raise SystemExit(f"unknown scene {name}")


# ---- reference arm ------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified reference compiled in
    place; falls back to the oracle port when that library did not travel), all host threads, rank 0 only."""
    if rank != 0:
        return
    from salviarenderer_b200 import abi as A, scenes
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so")
    kind = "reference"
    if not os.path.exists(ref_lib):
        import __graft_entry__ as g
        ref_lib, kind = g.build_oracle(), "port"
    be = A.Backend(ref_lib)
    what = None
    if args.scene == "sponza":
        sc = twin_scene(args, scenes, A)
    else:
        sc, what = small_scene(args.scene, scenes)
    sc.setup(be)
    budget_s = 150.0 if args.scene == "sponza" else 40.0
    t0 = time.perf_counter()
    sc.render(be, 0)
    be.flush()
    first = time.perf_counter() - t0
    warm = max(0, min(args.warmup, int(20.0 / max(first, 1e-3))) - 1)
    for i in range(warm):
        sc.render(be, (i + 1) % sc.n_frames)
    be.flush()
    steps = max(1, min(args.steps, int(budget_s / max(first, 1e-3))))
    be.query_begin()
    t0 = time.perf_counter()
    for i in range(steps):
        sc.render(be, i % sc.n_frames)
    be.flush()
    dt = time.perf_counter() - t0
    stats = be.query_get()
    fps = steps / dt
    cores = os.cpu_count() if kind == "reference" else 1
    if args.dump_frame:  # one frame's buffers for the product arm's `parity` (untimed)
        r = sc.run(be, args.parity_frame % sc.n_frames)
        np.savez(args.dump_frame, color=r.color, depth=r.depth, stencil=r.stencil,
                 resolved=r.resolved if r.resolved is not None else np.zeros(0, np.uint8),
                 stats=np.array([r.stats[k] for k in GATED], np.int64))
    cfg = config_dict(args, max(args.gpus, 1)) if args.scene == "sponza" else {"workload": what}
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "shaded_mpix_per_s": stats["ps_invocations"] / dt / 1e6,
        "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} full frames of the same workload (requested {args.steps})"
                                   + ("; the reference's SASL compiler cannot be built here: cpp twins of the SASL shaders on the "
                                      "unmodified reference core" if args.shaders == "sasl" and args.scene == "sponza" else "")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_subprocess(args, scene="sponza", frames=2, dump=None, parity_frame=0, timeout=900):
    """Times the reference CPU renderer on a bounded sample (and optionally dumps one frame), in a subprocess so that its thread
    pool and any crash stay out of this process."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(frames), "--warmup", "1",
           "--width", str(args.width), "--height", str(args.height), "--samples", str(args.samples), "--tex-size", str(args.tex_size),
           "--aniso", str(args.aniso), "--shaders", args.shaders, "--scene", scene, "--parity-frame", str(parity_frame)]
    if dump:
        cmd += ["--dump-frame", dump]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                cb = d["cpu_baseline"]
                cb["shaded_mpix_per_s"] = d.get("shaded_mpix_per_s")
                cb["ms_per_frame"] = d.get("ms_per_step")
                return cb
        return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": "failed: " + (out.stderr or out.stdout)[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}


def compare_with_dump(res, path, what):
    """Buffer-by-buffer comparison of a product frame with the reference's dump: the gates of SURVEY 8d."""
    ref = np.load(path)
    out = {"against": what, "bit_exact": True}
    n_px = res.color.shape[0] * res.color.shape[1]
    for k in ("depth", "stencil"):
        a, b = getattr(res, k), ref[k]
        a, b = (a.view(np.uint32), b.view(np.uint32)) if k == "depth" else (a, b)
        out[k + "_mismatches"] = int((a != b).sum())
    for k in ("color", "resolved"):
        a, b = getattr(res, k), ref[k]
        if a is None or b.size == 0:
            continue
        diff = np.abs(a.astype(np.int16) - b.astype(np.int16))
        out[k + "_max_abs_diff_lsb"] = int(diff.max())
        out[k + "_pixels_differing"] = int((diff > 0).any(axis=(-1, -2)).sum())
        out[k + "_pixels_differing_frac"] = out[k + "_pixels_differing"] / n_px
    out["counters"] = {k: [int(res.stats[k]), int(v)] for k, v in zip(GATED, ref["stats"])}
    out["counters_equal"] = all(a == b for a, b in out["counters"].values())
    out["bit_exact"] = (out["depth_mismatches"] == 0 and out["stencil_mismatches"] == 0 and out["counters_equal"]
                        and out.get("color_max_abs_diff_lsb", 0) == 0 and out.get("resolved_max_abs_diff_lsb", 0) == 0)
    out["within_north_star_tolerance"] = (out["depth_mismatches"] == 0 and out["stencil_mismatches"] == 0 and out["counters_equal"]
                                          and out.get("color_max_abs_diff_lsb", 0) <= 1 and out.get("color_pixels_differing_frac", 0) < 1e-4
                                          and out.get("resolved_max_abs_diff_lsb", 0) <= 1)
    return out


# ---- product arm ----------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import salviarenderer_b200 as pkg
    from salviarenderer_b200 import abi as A, scenes, sortfirst

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = world
    be = pkg.load(local)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream shared by the library, torch and NCCL
    torch.cuda.set_stream(stream)
    be.set_stream(stream.cuda_stream)  # kernels, copies and NCCL all order on torch's current stream

    def barrier():
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- the scene; SASL shaders are compiled by rank 0 first (the others then hit the on-disk cache) ----
    shaders_used, shader_note = args.shaders, None
    sc = scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=args.aniso)
    if args.shaders == "sasl":
        try:
            if rank == 0:
                install_sasl_shaders(sc, be, A)
            barrier()
            if rank != 0:
                install_sasl_shaders(sc, be, A)
        except Exception as e:  # noqa: BLE001 - e.g. no nvcc on the box: the built-in twins compute the same frames
            shaders_used = "builtin-twins"
            shader_note = f"SASL run-time compilation failed ({type(e).__name__}: {str(e)[:200]}); the built-in twins of the two shaders ran instead"
            sc = twin_scene(args, scenes, A)
    sc.setup(be)
    resolved = sc.t.resolved if sc.t.resolved is not None else sc.t.color

    # ---- sort-first frame assembly (N > 1): salviarenderer_b200/sortfirst.py.  Two frame buffers on rank 0: the ranks are not
    # in lockstep with rank 0's consumer, and the e2e readback of frame k overlaps frame k+1 ----
    targets = [resolved]
    if sc.t.resolved is not None:
        # frame buffers on rank 0 (SLV_BENCH_NBUF, default 2; N > 1: 4 - a rank may then run up to three frames ahead of the
        # slowest one, so per-frame load differences between the ranks average out instead of adding up in lockstep)
        nbuf = int(os.environ.get("SLV_BENCH_NBUF", "4" if n > 1 else "2"))
        for _ in range(max(nbuf, 2) - 1):
            targets.append(be.create_texture(args.width, args.height, 1, resolved.fmt))
    fg = sortfirst.FrameGather(be, targets, rank, n, "cuda")

    def frame(i):
        fg.begin_frame()
        if sc.t.resolved is not None:
            sc.t.resolved = fg.target()
        sc.render(be, i % sc.n_frames, before_resolve=fg.before_resolve)
        fg.gather()

    host_s = [0.0]

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        host_s[0] = time.perf_counter() - t0
        if finish is not None:
            finish()  # host-blocking: work the library put on its own streams (asynchronous readbacks) has completed
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if n > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- warm-up, then the headline: inputs resident in HBM ----
    for i in range(max(args.warmup, 3)):
        frame(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    be.query_begin()
    ms_total = timed(frame, args.steps)
    host_loop_us = host_s[0] / args.steps * 1e6
    launches = be.launch_count()
    stats = be.query_get()
    tr0 = be.traffic()
    clocks = sampler.stop() if rank == 0 else None
    cnt = torch.tensor([float(stats["ps_invocations"]), float(tr0["ps_executed"])], device="cuda", dtype=torch.float64)
    if n > 1:
        dist.all_reduce(cnt)
    ps_all, ps_exec_all = float(cnt[0].item()), float(cnt[1].item())
    ms_per_step = ms_total / args.steps
    fps = 1e3 / ms_per_step

    # ---- e2e: the frame's geometry comes from pinned host memory, the resolved frame goes back to the host ----
    vb_np, ib_np = sc.mesh.streams[0], sc.mesh.indices
    vb_host = torch.from_numpy(np.ascontiguousarray(vb_np)).pin_memory()
    ib_host = torch.from_numpy(np.ascontiguousarray(ib_np).view(np.int32)).pin_memory()
    vb_h, ib_h = sc.mesh.upload(be)[0][0], sc.mesh.upload(be)[1]
    out_bytes = args.width * args.height * 4
    h2d = vb_host.numel() * 4 + ib_host.numel() * 4
    d2h = out_bytes if rank == 0 else 0
    hf = None
    if sc.t.resolved is not None and n > 1:
        # N > 1: no frame is gathered on one GPU.  Every rank resolves its tiles locally and exports them straight into ONE pinned
        # host frame shared by the ranks (POSIX shared memory, slv_texture_export_tiles_async) over its OWN host link: the device ->
        # host traffic of a frame is spread over N links.  Two host frames / local targets alternate; the timed region ends after
        # every rank's last export has landed (slv_readback_wait on each rank, then the barrier of `timed`).
        hf = sortfirst.HostFrame(be, out_bytes, rank, n, nbuf=2)
        local_targets = [be.create_texture(args.width, args.height, 1, resolved.fmt) for _ in range(2)]
        d2h = sum(be.packed_tiles_bytes(resolved, r, n) for r in range(n))  # all ranks together: the whole frame
        # the replicated geometry: every rank uploads an N-th of it from pinned host memory, an NCCL all-gather over NVLink
        # assembles the vertex + index buffers on every GPU (sortfirst.ShardedUpload) - each input byte crosses a host link once
        # two buffer sets alternate, so the upload of frame k+1 does not wait for the geometry pass of frame k
        vb2 = be.create_buffer(np.ascontiguousarray(vb_np, dtype=np.float32))
        ib2 = be.create_buffer(np.ascontiguousarray(ib_np))
        geo_sets = [(vb_h, ib_h), (vb2, ib2)]
        su = sortfirst.ShardedUpload(be, [list(g) for g in geo_sets], [vb_np, ib_np], rank, n)
        h2d = su.h2d_bytes_per_rank * n
        kctr = [0]

        def frame_e2e(i):
            which = su.upload()
            for d in sc.frame_draws(be, i % sc.n_frames):  # this frame's draws read the set just uploaded
                d.streams[0].buffer, d.index_buffer = geo_sets[which]
            k = kctr[0]
            kctr[0] += 1
            sc.t.resolved = local_targets[k % 2]
            sc.render(be, i % sc.n_frames)
            hf.export(local_targets[k % 2], k)

        finish_e2e = be.readback_wait
        e2e_how = ("every rank: an N-th of the vertex+index data from pinned host memory to its GPU + an NCCL all-gather over NVLink into the "
                   "library's buffers (h2d_bytes_per_step is the sum over the ranks = the whole geometry once), the frame's draws on its own tiles, "
                   "MSAA resolve into a local surface, slv_texture_export_tiles_async of the tiles it owns into ONE pinned host frame shared "
                   "by the ranks (POSIX shared memory) over its own PCIe link, every step; two host frames alternate; d2h_bytes_per_step is "
                   "the sum over the ranks (= one whole frame); the timed region ends after every rank's last export has landed")
    elif sc.t.resolved is not None:
        # the assembled frame goes back through slv_texture_readback_async into one of two pinned buffers while the next frame
        # renders into the other frame buffer (what an application pipelining frames does); every step's upload and readback is
        # inside the timed region, which ends only after the last copy has landed
        out_host = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)] if rank == 0 else None

        def frame_e2e(i):
            be.upload_from_ptr(vb_h, vb_host.data_ptr(), vb_host.numel() * 4)
            be.upload_from_ptr(ib_h, ib_host.data_ptr(), ib_host.numel() * 4)
            tgt, slot = fg.target(), fg.frame % 2
            frame(i)
            if rank == 0:
                be.read_texture_into_async(tgt, out_host[slot].data_ptr(), out_bytes)

        finish_e2e = be.readback_wait
        e2e_how = ("slv_buffer_upload of the vertex+index buffers from pinned host memory, the full frame, and "
                   "slv_texture_readback_async of the resolved 4K frame into pinned host memory, every step; two frame buffers "
                   "/ host buffers alternate so that the readback of frame k overlaps the rendering of frame k+1; the timed region ends "
                   "after the last readback has landed (slv_readback_wait)")
    else:
        out_host1 = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()

        def frame_e2e(i):
            be.upload_from_ptr(vb_h, vb_host.data_ptr(), vb_host.numel() * 4)
            be.upload_from_ptr(ib_h, ib_host.data_ptr(), ib_host.numel() * 4)
            frame(i)
            if rank == 0:
                be.read_texture_into(resolved, out_host1.data_ptr(), out_bytes)  # synchronises: the app now owns the pixels

        finish_e2e = None
        e2e_how = ("slv_buffer_upload of the vertex+index buffers from pinned host memory, the full frame, and "
                   "slv_texture_readback of the frame into pinned host memory on rank 0, every step")

    for i in range(4):
        frame_e2e(i)
    if finish_e2e:
        finish_e2e()
    e2e_ms = timed(frame_e2e, args.steps, finish_e2e) / args.steps
    e2e_host_us = host_s[0] / args.steps * 1e6  # host time of the enqueue loop alone (Python + torch + NCCL enqueue + ctypes)
    if hf is not None:
        sc.t.resolved = resolved
        for f in range(sc.n_frames):
            for d in sc.frame_draws(be, f):
                d.streams[0].buffer, d.index_buffer = geo_sets[0]

    # ---- roofline: per-stage CUDA events on the launching stream, algorithmic bytes from exact counters ----
    be.profile_enable(True)
    be.query_begin()
    for i in range(args.steps):
        frame(i)
    be.flush()
    stages = be.profile_stages()
    traffic = be.traffic()
    st2 = be.query_get()
    be.profile_enable(False)
    K = args.steps
    stage_ms = {"geometry": stages["geometry"] / K, "scan+bin_fill": stages["bin"] / K, "sort": stages["sort"] / K,
                "region_bin": stages["region_bin"] / K, "cover": stages["raster_or_cover"] / K, "shade": stages["shade"] / K}
    per_rank = None
    if n > 1:  # which stage limits each rank (sort-first: the front half is replicated, the back half is sharded)
        t = torch.tensor([list(stage_ms.values())], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(n)]
        dist.all_gather(allt, t)
        per_rank = [{k: float(v) for k, v in zip(stage_ms, x[0].tolist())} for x in allt]
        tsum = torch.tensor([traffic["z_tested"], traffic["z_written"], traffic["c_written"], traffic["c_read"],
                             st2["cprimitives"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tsum)
        traffic = dict(traffic, z_tested=float(tsum[0]), z_written=float(tsum[1]), c_written=float(tsum[2]), c_read=float(tsum[3]))
        st2 = dict(st2, cprimitives=float(tsum[4]) / n)  # replicated on every rank
    cover_ms, shade_ms = stage_ms["cover"], stage_ms["shade"]
    R = 5  # position + 4 attributes (VS_SPONZA)
    S, W, H = args.samples, args.width, args.height
    vstride, n_prims = 48, 262249
    n_unique = len(np.unique(ib_np))
    # B_alg = B_clear + B_geom + B_frag + B_tex + B_resolve (per frame)
    b_clear = W * H * S * (4 + 8)
    b_geom = n_prims * 3 * 4 + n_unique * vstride + n_unique * 16 * R
    b_setup = st2["cprimitives"] / K * 3 * 16 * R
    b_depth = (8 * traffic["z_tested"] + 8 * traffic["z_written"]) / K
    b_color = (4 * traffic["c_written"] + 4 * traffic["c_read"]) / K
    # B_tex: the texture bytes a frame TOUCHES - every mip level of every bound texture that at least one sampler call
    # of the frame reads, once.  Counted on the device in an untimed pass (slv_texture_level_tracking: a bit mask of sampled levels
    # per texture), per distinct frame of the animation, union over the ranks, average over the frames.
    b_tex_full = 24 * sum(max(args.tex_size >> l, 1) ** 2 * 4 for l in range(args.tex_size.bit_length()))
    per_frame = []
    for f in range(sc.n_frames):
        be.texture_level_tracking(True)
        frame(f)
        be.flush()
        masks = [be.texture_levels_touched(t) for t in sc.textures]
        if n > 1:  # union over the ranks (NCCL has no bitwise OR: one 0 / 1 entry per (texture, level), MAX)
            bits = torch.tensor([[(m >> l) & 1 for l in range(16)] for m in masks], device="cuda", dtype=torch.int32)
            dist.all_reduce(bits, op=dist.ReduceOp.MAX)
            masks = [sum(int(b) << l for l, b in enumerate(row)) for row in bits.tolist()]
        touched = 0
        for t, m in zip(sc.textures, masks):
            for l in range(be.level_count(t)):
                if m & (1 << l):
                    lw, lh = be.level_size(t, l)
                    touched += lw * lh * 4
        per_frame.append(touched)
    be.texture_level_tracking(False)
    b_tex = sum(per_frame) / len(per_frame)
    b_tex_how = (f"mip levels of the 24 bound textures that the frame's sampler calls read (device-side level masks, average over the "
                 f"{sc.n_frames} distinct frames: {b_tex / 1e6:.1f} MB of the {b_tex_full / 1e6:.1f} MB the full chains hold)")
    b_resolve = W * H * 4 * (S + 1) if S > 1 else 0
    b_frame = b_clear + b_geom + b_setup + b_depth + b_color + b_tex + b_resolve
    peak, peak_src = peaks()
    shade_kernel = f"slv_jit_k_shade_s{S} (shade_quad_main<{S}, SASL pixel shader>)" if shaders_used == "sasl" else (
        f"k_shade<{S}, PS_SPONZA_GRAD>" if shaders_used == "builtin-twins" else f"k_shade<{S}, PS_SPONZA>")
    kernels = {  # algorithmic bytes per launch (one launch per frame; N > 1: summed over the ranks' launches, slowest rank's time)
        "k_cover": {"name": f"k_cover<{S}>", "bytes": b_depth + b_setup, "ms": cover_ms,
                    "what": "8 B per depth-tested sample + 8 B per depth-written sample + one read of every emitted triangle's setup"},
        "k_shade": {"name": shade_kernel, "bytes": b_color + b_tex, "ms": shade_ms,
                    "what": "4 B per colour-written sample (+4 B per blended read) + texture bytes: " + b_tex_how},
    }
    if per_rank:
        kernels["k_cover"]["ms"] = max(r["cover"] for r in per_rank)
        kernels["k_shade"]["ms"] = max(r["shade"] for r in per_rank)
    for v in kernels.values():
        v["achieved_gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        v["frac"] = v["achieved_gbs"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    traffic_measured = None
    tpath = os.path.join(ROOT, "profiles", "dram_bytes.json")
    if os.path.exists(tpath):
        traffic_measured = json.load(open(tpath)).get(kernels[dom]["name"].split(" ")[0].split("<")[0], {}).get("dram_bytes_per_launch")
    stage_sum = sum(stage_ms.values())
    roofline = {"bound": "hbm", "kernel": kernels[dom]["name"],
                "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": traffic_measured, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["bytes"], "launches_per_frame": 1,
                "kernel_ms_per_launch": kernels[dom]["ms"], "kernel_share_of_step": kernels[dom]["ms"] / max(stage_sum, 1e-9),
                "kernels": kernels,
                "stage_ms_per_frame": stage_ms,
                "frame": {"algorithmic_bytes": b_frame, "achieved_gbs": b_frame / (ms_per_step * 1e-3) / 1e9,
                          "frac": b_frame / (ms_per_step * 1e-3) / 1e9 / peak,
                          "terms": {"clear": b_clear, "geometry": b_geom, "setup": b_setup, "depth": b_depth, "colour": b_color,
                                    "texture": b_tex, "resolve": b_resolve}},
                "note": "kernel times from CUDA events around each kernel (slv_profile_get_stages) in a separate pass over the same "
                        "K frames; `frame` divides the whole frame's algorithmic bytes (SURVEY 8d) by the headline ms_per_step"}
    if per_rank:
        roofline["stage_ms_per_frame_per_rank"] = per_rank

    # ---- parity (N == 1): one full-size frame of the product against the unmodified reference renderer, untimed; the same
    # subprocess times the reference on a bounded sample (cpu_baseline) ----
    cpu_baseline, parity = None, None
    if rank == 0 and n == 1 and not (args.no_cpu_baseline and args.no_parity):
        dump = None if args.no_parity else os.path.join(tempfile.mkdtemp(prefix="slv_parity_"), "ref_frame.npz")
        cpu_baseline = run_reference_subprocess(args, "sponza", args.cpu_baseline_frames, dump, args.parity_frame)
        if dump and os.path.exists(dump):
            res = sc.run(be, args.parity_frame % sc.n_frames)
            parity = compare_with_dump(res, dump, "oracle/_ref (the unmodified reference renderer; cpp twins of the SASL shaders), "
                                       f"frame {args.parity_frame % sc.n_frames} at {W}x{H}x{S}, every sample of colour / depth / stencil, "
                                       "the resolved frame, six counters")
            os.remove(dump)
        elif dump:
            parity = {"against": "oracle/_ref", "error": "the reference subprocess produced no frame: " + str(cpu_baseline.get("sample"))[:200]}

    # ---- variants (N == 1): the same frames with other shaders / filters ----
    variants = None
    if n == 1 and not args.no_variants:
        variants = run_variants(args, be, timed, scenes, A)
        try:  # frame-level roofline of each variant: the same frames, so the same algorithmic bytes but for the texture term, which
            # is charged as the headline's (the level masks were counted on the headline's sampler state)
            for v in variants.values():
                if isinstance(v, dict) and v.get("ms_per_step"):
                    v["frame_roofline"] = {"algorithmic_bytes": b_frame, "achieved_gbs": b_frame / (v["ms_per_step"] * 1e-3) / 1e9,
                                           "frac": b_frame / (v["ms_per_step"] * 1e-3) / 1e9 / peak, "peak": peak, "unit": "GB/s"}
        except Exception:  # noqa: BLE001 - reporting only
            pass
    other = None
    if n == 1 and not args.no_configs:
        other = run_other_configs(args, be, scenes, A)

    if rank == 0:
        line = {
            "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": n, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "shaded_mpix_per_s": ps_all / (ms_total * 1e-3) / 1e6,
            "shaded_mpix_per_s_executed": ps_exec_all / (ms_total * 1e-3) / 1e6,
            "ps_invocations_per_frame": ps_all / args.steps,
            "ps_lanes_executed_per_frame": ps_exec_all / args.steps,
            "shaded_note": "ps_invocations = the reference's counter (pixels of every quad with a live sample after early-Z, what the "
                           "reference shades); executed = shader lanes the visibility-first path actually ran (one per pixel / quad lane "
                           "and distinct final owner)",
            "config": config_dict(args, n),
            "shaders_run": shaders_used, "shader_note": shader_note,
            "sortfirst_transport": fg.transport,
            "clocks": clocks,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "host_loop_us_per_frame": e2e_host_us,
                    "what": e2e_how},
            "gpu_launches": int(launches),
            "host_frame_loop_us_per_frame": host_loop_us,
            "roofline": roofline,
            "parity": parity,
            "variants": variants,
            "configs": other,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if n > 1:
        if hf is not None:
            hf.close()
        fg.close()
        dist.destroy_process_group()


def run_variants(args, be, timed, scenes, A):
    """The workload with other shader / filter combinations, reported beside the headline, never instead of it:
    samples/Sponza's own cpp shaders with trilinear samplers (round 1's headline), the SASL pair with trilinear samplers, and
    the built-in twins of the SASL pair with the headline's samplers (pixel-granular k_shade)."""
    out = {}
    steps = max(10, min(args.steps, 40))

    def run(key, sc, what):
        sc.setup(be)
        for i in range(3):
            sc.render(be, i)
        be.query_begin()
        ms = timed(lambda i: sc.render(be, i % sc.n_frames), steps) / steps
        st, tr = be.query_get(), be.traffic()
        out[key] = {"frames_per_sec": 1e3 / ms, "ms_per_step": ms, "steps": steps, "what": what,
                    "ps_lanes_executed_per_frame": tr["ps_executed"] / steps, "ps_invocations_per_frame": st["ps_invocations"] / steps}

    try:
        if not (args.shaders == "builtin" and args.aniso <= 1):
            run("builtin_cpp_shaders_trilinear", scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size),
                "SLV_VS_SPONZA + SLV_PS_SPONZA (samples/Sponza's cpp shaders, tex2d: LOD once per quad), trilinear - round 1's headline")
        if not (args.shaders == "sasl" and args.aniso <= 1):
            sc = scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size)
            install_sasl_shaders(sc, be, A)
            run("sasl_vs_ps_trilinear", sc, "SASL vertex + pixel shader, trilinear samplers")
        if args.shaders == "sasl":
            run("builtin_twins_of_the_sasl_pair", twin_scene(args, scenes, A),
                "SLV_VS_SPONZA + SLV_PS_SPONZA_GRAD with the headline's samplers: the same frames through the pixel-granular k_shade")
    except Exception as e:  # noqa: BLE001 - e.g. no nvcc on the box: the variants are optional
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def run_other_configs(args, be, scenes, A):
    """BASELINE.json configs[0], [1], [2], [4] at their full sizes: frames/s of the product (CUDA events, inputs resident) and of
    the reference CPU renderer on the same box (bounded sample, subprocess).  Parity of each is pinned by tests/ (-m gpu)."""
    out = {}
    for key, steps, ref_frames in (("c1", 100, 20), ("c2", 100, 10), ("c3a", 100, 10), ("c3b", 50, 7), ("c5", 5, 1)):
        try:
            t0 = time.perf_counter()
            sc, what = small_scene(key, scenes)
            sc.setup(be)
            for i in range(2):
                sc.render(be, i % sc.n_frames)
            be.flush()
            be.query_begin()
            be.event_record(0)
            for i in range(steps):
                sc.render(be, i % sc.n_frames)
            be.event_record(1)
            ms = be.event_elapsed_ms(0, 1) / steps
            st = be.query_get()
            out[key] = {"what": what, "frames_per_sec": 1e3 / ms, "ms_per_step": ms, "steps": steps,
                        "shaded_mpix_per_s": st["ps_invocations"] / steps / ms / 1e3,
                        "ia_primitives_per_frame": st["ia_primitives"] / steps}
            if not args.no_cpu_baseline:
                cb = run_reference_subprocess(args, key, ref_frames, timeout=600)
                out[key]["cpu_baseline"] = {k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample", "ms_per_frame")}
            out[key]["wall_s"] = time.perf_counter() - t0
            del sc
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — frames/s (and shaded Mpix/s) of the Sponza-class scene at 3840x2160, 4x MSAA (BASELINE.json
configs[3]) on N B200s, through the C ABI of the CUDA product.

    python bench.py --gpus N --steps K --warmup W            # the product (torchrun launches N ranks for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU renderer (rank 0 only)

A "step" is one frame: clear colour + clear depth/stencil + 24 draws (one per material group) + MSAA resolve
(+ for N>1 the sort-first gather of the owned tiles to rank 0 over NCCL).  All inputs are resident in HBM when the
timed region starts (`value`); `e2e` repeats the measurement with the frame's geometry uploaded from pinned host
memory and the resolved frame read back to pinned host memory inside the timed region, every step.

Timing: CUDA events recorded on the stream the kernels run on, barrier + synchronize on both sides, max over
ranks.  Inputs are larger than L2 (the 4K 4xMSAA colour + depth/stencil targets alone are 398 MB vs 126 MB of L2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "Sponza-like atrium (262,249 tris, 24 material draws, trilinear), 3840x2160, 4x MSAA + resolve"


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
    ap.add_argument("--aniso", type=int, default=0, help="max anisotropy (0 = trilinear, as samples/Sponza)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the SASL-shader variants of the workload")
    ap.add_argument("--cpu-baseline-frames", type=int, default=2)
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


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified reference compiled in
    place; falls back to the oracle port when that library did not travel), all host threads, rank 0 only."""
    if rank != 0:
        return
    from salviarenderer_b200 import abi, scenes
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so")
    kind = "reference"
    if not os.path.exists(ref_lib):
        import __graft_entry__ as g
        ref_lib, kind = g.build_oracle(), "port"
    be = abi.Backend(ref_lib)
    sc = scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=args.aniso)
    sc.setup(be)
    budget_s = 150.0
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
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "shaded_mpix_per_s": stats["ps_invocations"] / dt / 1e6,
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} full frames of the same workload (requested {args.steps})"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, n):
    return {"workload": WORKLOAD if (args.width, args.height, args.samples) == (3840, 2160, 4) else
            f"Sponza-like atrium {args.width}x{args.height}x{args.samples}",
            "width": args.width, "height": args.height, "msaa": args.samples, "triangles": 262249, "draws_per_frame": 24,
            "texture": f"24 x {args.tex_size}^2 rgba8 + mips, wrap, " + (f"{args.aniso}x anisotropic" if args.aniso > 1 else "trilinear"),
            "color_format": "bgra8", "depth_stencil_format": "rg32f",
            "parallelism": "single GPU" if n == 1 else f"sort-first: 64x64 screen tiles interleaved over {n} GPUs, geometry replicated, "
                           "finished tiles resolved straight into rank 0's surface over NVLink peer memory (fallback: NCCL gather)",
            "l2_policy": "inputs larger than L2 (398 MB of render targets per frame vs 126 MB L2); no explicit flush"}


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
    from salviarenderer_b200 import scenes, sortfirst

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
    sc = scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=args.aniso)
    sc.setup(be)
    resolved = sc.t.resolved if sc.t.resolved is not None else sc.t.color

    # ---- sort-first frame assembly (N > 1): salviarenderer_b200/sortfirst.py.  Two frame buffers on rank 0: the ranks are not
    # in lockstep with rank 0's consumer, and the e2e readback of frame k overlaps frame k+1 ----
    targets = [resolved]
    if sc.t.resolved is not None:
        targets.append(be.create_texture(args.width, args.height, 1, resolved.fmt))
    fg = sortfirst.FrameGather(be, targets, rank, n, "cuda")

    def frame(i):
        fg.begin_frame()
        if sc.t.resolved is not None:
            sc.t.resolved = fg.target()
        sc.render(be, i % sc.n_frames, before_resolve=fg.before_resolve)
        fg.gather()

    def barrier():
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
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
    launches = be.launch_count()
    stats = be.query_get()
    clocks = sampler.stop() if rank == 0 else None
    ps_all = torch.tensor([float(stats["ps_invocations"])], device="cuda", dtype=torch.float64)
    if n > 1:
        dist.all_reduce(ps_all)
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
    if sc.t.resolved is not None:
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
                   "slv_texture_readback_async of the resolved 4K frame into pinned host memory (rank 0), every step; two frame buffers "
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
    cover_ms, shade_ms = stages["raster_or_cover"] / K, stages["shade"] / K
    geom_ms, bin_ms, sort_ms, rbin_ms = stages["geometry"] / K, stages["bin"] / K, stages["sort"] / K, stages["region_bin"] / K
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
    b_tex = 24 * sum(max(args.tex_size >> l, 1) ** 2 * 4 for l in range(args.tex_size.bit_length()))
    b_resolve = W * H * 4 * (S + 1) if S > 1 else 0
    b_frame = b_clear + b_geom + b_setup + b_depth + b_color + b_tex + b_resolve
    peak, peak_src = peaks()
    kernels = {  # algorithmic bytes per launch (one launch per frame) and measured launch time
        "k_cover": {"bytes": b_depth + b_setup, "ms": cover_ms,
                    "what": "8 B per depth-tested sample + 8 B per depth-written sample + one read of every emitted triangle's setup"},
        "k_shade": {"bytes": b_color + b_tex, "ms": shade_ms,
                    "what": "4 B per colour-written sample (+4 B per blended read) + the bound textures' mip chains once"},
    }
    for v in kernels.values():
        v["achieved_gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        v["frac"] = v["achieved_gbs"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    traffic_measured = None
    tpath = os.path.join(ROOT, "profiles", "dram_bytes.json")
    if os.path.exists(tpath):
        traffic_measured = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
    stage_sum = geom_ms + bin_ms + sort_ms + rbin_ms + cover_ms + shade_ms
    roofline = {"bound": "hbm", "kernel": dom + f"<{S}>" if dom == "k_cover" else dom + f"<{S}, PS_SPONZA>",
                "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": traffic_measured, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["bytes"], "launches_per_frame": 1,
                "kernel_ms_per_launch": kernels[dom]["ms"], "kernel_share_of_step": kernels[dom]["ms"] / max(stage_sum, 1e-9),
                "kernels": kernels,
                "stage_ms_per_frame": {"geometry": geom_ms, "scan+bin_fill": bin_ms, "sort": sort_ms, "region_bin": rbin_ms, "cover": cover_ms, "shade": shade_ms},
                "frame": {"algorithmic_bytes": b_frame, "achieved_gbs": b_frame / (ms_per_step * 1e-3) / 1e9,
                          "frac": b_frame / (ms_per_step * 1e-3) / 1e9 / peak,
                          "terms": {"clear": b_clear, "geometry": b_geom, "setup": b_setup, "depth": b_depth, "colour": b_color,
                                    "texture": b_tex, "resolve": b_resolve}},
                "note": "kernel times from CUDA events around each kernel (slv_profile_get_stages) in a separate pass over the same "
                        "K frames; `frame` divides the whole frame's algorithmic bytes (SURVEY 8d) by the headline ms_per_step"}

    # ---- variants (N == 1): the same frames with SASL shaders compiled at run time ----
    variants = None
    if n == 1 and not args.no_variants:
        variants = run_variants(args, be, sc, timed)

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample (rank 0, N == 1) ----
    cpu_baseline = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(args)

    if rank == 0:
        line = {
            "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": n, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "shaded_mpix_per_s": float(ps_all.item()) / (ms_total * 1e-3) / 1e6,
            "ps_invocations_per_frame": float(ps_all.item()) / args.steps,
            "config": dict(config_dict(args, n), sortfirst_transport=fg.transport),
            "clocks": clocks,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "frames/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "what": e2e_how},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "variants": variants,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if n > 1:
        fg.close()
        dist.destroy_process_group()


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


def run_variants(args, be, sc, timed):
    """The workload with SASL shaders compiled at run time (salviarenderer_b200/sasl): (1) the sample's SASL vertex shader in
    place of its cpp twin — samples/Sponza runs exactly this pair (SASL VS + cpp PS); (2) additionally a SASL pixel shader
    (tex2D = sample_2d_grad), with the headline's trilinear samplers and with 16x anisotropic ones — the only way the
    reference can filter anisotropically; SASL pixel shaders run in the module's quad-granular k_shade on
    the visibility-first path.  Reported beside the headline, never instead of it."""
    import numpy as np
    from salviarenderer_b200 import abi as A, scenes
    out = {}
    try:
        from salviarenderer_b200.sasl import jit
        steps = max(10, min(args.steps, 60))
        vs = jit.compile(SASL_VS_SPONZA, "vs")
        vs_mod = jit.load(be, vs)

        def vs_binding(wvp, light, eye):
            return A.shader_binding(A.program_jit(vs_mod), vs.unit.pack_uniforms(
                {"wvpMatrix": np.asarray(wvp, np.float32).reshape(4, 4), "lightPos": light, "eyePos": eye}))

        sc.vs_binding = vs_binding
        sc._draw_cache = {}
        for i in range(3):
            sc.render(be, i)
        ms = timed(lambda i: sc.render(be, i % sc.n_frames), steps) / steps
        out["sasl_vertex_shader"] = {"frames_per_sec": 1e3 / ms, "ms_per_step": ms, "steps": steps,
                                     "what": "SASL vertex shader (JIT module) + built-in pixel shader, trilinear"}
        sc.vs_binding = None
        sc._draw_cache = {}

        ps = jit.compile(SASL_PS_SPONZA, "ps")
        ps_mod = jit.load(be, ps)
        # BASELINE configs[3] literally: the headline scene with SASL vertex AND pixel shaders (trilinear), then with 16x
        # anisotropic samplers; both on the visibility-first path (k_cover + the module's quad-granular k_shade)
        for key, aniso, what in (
                ("sasl_vs_ps", args.aniso, "SASL vertex + pixel shader (tex2D with per-pixel derivatives), the headline's samplers, "
                                           "visibility-first path (quad-granular k_shade)"),
                ("sasl_vs_ps_aniso16", 16, "SASL vertex + pixel shader (tex2D with per-pixel derivatives), 16x anisotropic samplers, "
                                           "visibility-first path (quad-granular k_shade)")):
            sc2 = scenes.SponzaLike(args.width, args.height, args.samples, tex_size=args.tex_size, max_aniso=aniso)
            sc2.setup(be)
            sc2.vs_binding = vs_binding
            draws = {}

            def render2(i, sc2=sc2, draws=draws):
                f = i % sc2.n_frames
                if f not in draws:
                    ds = sc2.frame_draws(be, f)
                    for d, (m, _, _) in zip(ds, sc2.groups):
                        d.ps = A.shader_binding(A.program_jit(ps_mod), b"", [sc2.samplers[m]])
                    draws[f] = ds
                sc2.render(be, f)

            for i in range(3):
                render2(i)
            steps2 = max(5, steps // 2)
            be.query_begin()
            ms2 = timed(render2, steps2) / steps2
            st = be.query_get()
            tr = be.traffic()
            out[key] = {"frames_per_sec": 1e3 / ms2, "ms_per_step": ms2, "steps": steps2, "what": what,
                        "ps_lanes_executed_per_frame": tr["ps_executed"] / steps2,
                        "ps_invocations_per_frame": st["ps_invocations"] / steps2}
    except Exception as e:  # noqa: BLE001 - e.g. no nvcc on the box: the variants are optional
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def run_cpu_baseline(args):
    """Times the reference CPU renderer on a bounded sample, in a subprocess (so that its thread pool and any
    crash stay out of this process)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_baseline_frames),
           "--warmup", "1", "--width", str(args.width), "--height", str(args.height), "--samples", str(args.samples),
           "--tex-size", str(args.tex_size), "--aniso", str(args.aniso)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
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


if __name__ == "__main__":
    main()

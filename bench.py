#!/usr/bin/env python
"""Headline benchmark: 30-s chunks/sec, audio -> piano-roll (-> notes), CNNRNNModelLarge.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chunks C] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path (log-mel -> forward -> sigmoid -> threshold ->
note grouping) over one batch of C synthetic 30-s chunks PER GPU (weak scaling: chunks are
independent units, each rank owns its block, no data-path collective).

  value : chunks/s over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e   : same metric through the public API with HOST buffers (pipeline.StreamingTranscriber): pinned wav
          -> H2D -> path -> D2H of the binary piano-rolls and the note list, every step, inside the timed
          region; the copies of neighbouring steps overlap the compute (3 streams, 2 buffer slots)
  roofline     : the dominant tensor kernel (by device time; per-stage CUDA events in a second pass of K steps)
  cpu_baseline : the oracle port (reference algorithm on the host cores), bounded sample

--impl reference times that CPU path alone (the reference is pure Python/PyTorch + librosa;
it cannot be pip-installed offline -- see DESIGN.md -- so the arm runs the oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MELS, HIDDEN, LAYERS = 320, 512, 3
N_SAMPLES, T_FRAMES = 480000, 938
METRIC, UNIT = "chunks_per_sec_audio_to_pianoroll", "30s-chunks/s"


# ------------------------------------------------------------------ algorithmic work per stage
def stage_flops(B: int) -> dict:
    """Algorithmic (unpadded) FLOPs of the GEMM-class stages for B chunks."""
    T = T_FRAMES
    F1, F2, F3 = 160, 80, 40
    H, Hl, D = HIDDEN, HIDDEN // 2, 3 * HIDDEN
    conv = lambda F, co, ci, k: 2.0 * F * T * co * ci * k
    f = {
        "res1.c1": conv(F1, 64, 32, 9),
        "res1.c2": conv(F1, 64, 64, 9) + conv(F1, 64, 32, 1),
        "res2.c1": conv(F2, 128, 64, 9),
        "res2.c2": conv(F2, 128, 128, 9) + conv(F2, 128, 64, 1),
        "freq": conv(F2, 256, 128, 21),
        "rnn0.gemm": 2.0 * T * (8 * H + 8 * Hl) * 256 * F3,
        "rnn1.gemm": 2.0 * T * 8 * H * 2 * H,
        "rnn2.gemm": 2.0 * T * 8 * H * 2 * H,
        "rnn0.rec": 2.0 * T * (2 * 4 * H * H + 2 * 4 * Hl * Hl),
        "rnn1.rec": 2.0 * T * 2 * 4 * H * H,
        "rnn2.rec": 2.0 * T * 2 * 4 * H * H,
        "attn.qkv": 2.0 * T * 3 * D * D,
        "attn.core": 2.0 * 2 * T * T * D,
        "attn.proj": 2.0 * T * D * D,
        "fc1": 2.0 * T * H * D,
        "heads": 2.0 * T * 3 * 88 * H,
    }
    return {k: v * B for k, v in f.items()}


def ncu_traffic(stage: str, chunks: int):
    """DRAM bytes of one launch of `stage` from the committed ncu --set full capture (same chunk count only)."""
    p = os.path.join(ROOT, "profiles", "r1_traffic_chunks64.json")
    try:
        d = json.load(open(p))
        return d["dram_bytes_per_launch"].get(stage) if d.get("chunks") == chunks else None
    except Exception:
        return None


def measured_burst():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(p)).get("bf16_tflops")
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, sampled every few ms through NVML while the timed region runs
    (nvidia-smi, one process per sample, is the fallback: a 100-ms region would only get one sample)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_mhz, self.reasons, self._stop_evt = index, [], None, set(), threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        if len(f) >= 6 and f[0].replace(".", "").isdigit():
            self.sm.append(float(f[0]))
            self.max_mhz = float(f[1]) if f[1].replace(".", "").isdigit() else self.max_mhz
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.004 if self.nvml else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------ CPU reference path (oracle port)
def cpu_reference_chunks_per_sec(n_chunks: int, threads: int, repeat: int = 1):
    """The reference's own B=1 loop (main.py:258-275) restated in oracle/: log-mel -> fp32 forward ->
    sigmoid -> threshold -> grouping, on the host cores.  Returns (chunks/s, seconds)."""
    from music_transcription_b200 import synth
    from oracle import frontend as ofe, model as omodel, notes as onotes
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict("cnn_rnn_large", N_MELS, HIDDEN, LAYERS, seed=1, gain=3 ** -0.5)
    fb = ofe.mel_filterbank(n_mels=N_MELS)
    waves = [synth.piano_chord(k) for k in range(n_chunks)]
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        rolls = []
        for y in waves:
            mel = torch.from_numpy(ofe.logmel(y, n_mels=N_MELS, fb=fb))[None, None]
            logits = omodel.large_forward(sd, mel, HIDDEN, LAYERS)
            rolls.append(onotes.threshold_roll(torch.sigmoid(logits)[0].numpy(), 0.5))
        onotes.group_notes(onotes.combine_piano_rolls(rolls))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_chunks / best, best


def gpu_eager_chunks_per_sec(wav: torch.Tensor, n_chunks: int, reps: int = 3):
    """SURVEY.md 8(d) 'GPU baseline beside it': the reference's modules as plain torch fp32 eager on this GPU
    (oracle port: F.conv2d / nn.LSTM / matmul -> cuDNN and cuBLAS kernels, TF32 off as in the reference), fed the
    log-mel of our frontend, batched like our step.  A reported baseline; nothing of it is on the product path."""
    from music_transcription_b200 import pipeline, synth
    from oracle import model as omodel
    dev = wav.device
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = {k: v.to(dev) for k, v in synth.synth_state_dict("cnn_rnn_large", N_MELS, HIDDEN, LAYERS, seed=1, gain=3 ** -0.5).items()}
        mel = pipeline.Frontend.get(device=dev).logmel(wav[:n_chunks]).clone()
        best = None
        for _ in range(reps + 1):                              # first pass = warm-up (cuDNN autotune, allocator)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            logits = omodel.large_forward(sd, mel, HIDDEN, LAYERS)
            roll = (torch.sigmoid(logits) > 0.5).float()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        del roll
        return n_chunks / (best / 1e3), best
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_chunks
    for _ in range(args.warmup):
        cpu_reference_chunks_per_sec(1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_chunks_per_sec(n, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args.chunks, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} chunk(s)/step x {args.steps} steps, B=1 loop as main.py:258-266, oracle port "
                                       "(numpy log-mel + fp32 torch forward + numpy grouping), all host threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(chunks, gpus):
    return {"workload": "BASELINE configs[3]-style: CNNRNNModelLarge (89M, n_mels 320, hidden 512, 3 layers) audio->piano-roll->notes, "
                        f"{chunks} synthetic 30-s chunks per GPU per step (weak scaling of the 240-chunk recording)",
            "chunks_per_gpu": chunks, "global_chunks": chunks * gpus, "samples_per_chunk": N_SAMPLES, "frames": T_FRAMES,
            "threshold": 0.5, "parallelism": f"chunk-sharded x{gpus}",
            "l2": "inputs 1.9 MB/chunk + ~200 MB/chunk of streamed activations per step >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------ ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunks", type=int, default=64, help="30-s chunks per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-chunks", type=int, default=2, help="chunks per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-chunks", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-eager-baseline", type=int, default=0, metavar="CHUNKS",
                    help="also time the oracle port (plain torch fp32 eager: cuDNN / cuBLAS, TF32 off) on this GPU for a batch "
                         "of CHUNKS chunks -- SURVEY 8d 'GPU baseline beside it'; adds gpu_eager_baseline to the JSON line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
        args.gpus = world

    import torch.distributed as dist
    from music_transcription_b200 import _lib, pipeline, sharding, synth
    from music_transcription_b200.transcription_model import TranscriptionModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().amt_device_check())

    C = args.chunks
    # synthetic waveforms: 8 distinct chord chunks tiled to C with a per-chunk gain (cheap to build)
    base = synth.cheap_wave_batch(8, N_SAMPLES, seed=rank)
    host_wav = torch.empty(C, N_SAMPLES, dtype=torch.float32).pin_memory()
    for i in range(C):
        host_wav[i] = base[i % 8] * (1.0 - 0.01 * (i // 8))
    wav = host_wav.to(dev)

    model = TranscriptionModel("cnn_rnn_large", n_mels=N_MELS, hidden_size=HIDDEN, num_layers=LAYERS, dropout=0.2, device=dev)
    model.load_state_dict(synth.synth_state_dict("cnn_rnn_large", N_MELS, HIDDEN, LAYERS, seed=1, gain=3 ** -0.5))
    model.eval()
    fe = pipeline.Frontend.get(device=dev)
    L = _lib.lib()
    cap = 88 * ((C * T_FRAMES + 1) // 2)
    probs = torch.empty(C, 88, T_FRAMES, device=dev)
    roll = torch.empty(C, 88, T_FRAMES, device=dev)
    notes = torch.empty(cap, 3, dtype=torch.int32, device=dev)
    counts = torch.empty(89, dtype=torch.int32, device=dev)
    stream = _lib.stream_ptr(dev)

    def step_device(w):
        mel = fe.logmel(w)
        logits = model(mel)
        _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.5, _lib.ptr(probs), _lib.ptr(roll), stream))
        pipeline.extract_notes_async(probs, 0.5, notes, counts)

    # End to end through the public API with HOST buffers: pipeline.StreamingTranscriber copies batch i+1 in and
    # batch i-1 out while batch i computes; every step still moves its own audio in and its rolls + notes out.
    streamer = pipeline.StreamingTranscriber(model, C, N_SAMPLES, 0.5)
    e2e_notes = [0]

    def run_e2e(steps):
        for _, nts in streamer.run(host_wav for _ in range(steps)):
            e2e_notes[0] = len(nts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 1)):
        step_device(wav)
    torch.cuda.synchronize()

    # ---- end to end through the public API with host buffers (the headline; measured first, right after warm-up)
    run_e2e(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler_e2e = ClockSampler(local_rank)
    sampler_e2e.start()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    clocks_e2e = sampler_e2e.stop()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    n_notes = e2e_notes[0]
    e2e_value = C * world * args.steps / (ms_e2e / 1e3)
    h2d = host_wav.numel() * 4
    d2h = streamer.roll_bytes + n_notes * 12

    # ---- timed region (device-resident inputs): exactly K steps, clocks sampled, no per-stage events
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.amt_launch_count()
    ms = timed(lambda: step_device(wav), args.steps)
    launches = int(L.amt_launch_count() - launches0)
    clocks = sampler.stop()
    value = C * world * args.steps / (ms / 1e3)

    # ---- the same K steps again with CUDA events around every kernel launch: the per-stage breakdown
    model.profile(True)
    timed(lambda: step_device(wav), args.steps)
    stages = model.profile_read()
    model.profile(False)

    # ---- the stages outside amt_model_forward, timed alone (same stream, CUDA events)
    mel_keep = fe.logmel(wav)
    logits_keep = model(mel_keep)

    def post():
        _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits_keep), logits_keep.numel(), 0.5, _lib.ptr(probs), _lib.ptr(roll), stream))
        pipeline.extract_notes_async(probs, 0.5, notes, counts)
    ms_logmel = timed(lambda: fe.logmel(wav), args.steps) / args.steps
    ms_post = timed(post, args.steps) / args.steps
    del mel_keep, logits_keep

    # ---- multi-GPU: the one collective of the path (note lists), outside the steady-state loop
    gathered = None
    if world > 1:
        local = notes[:int(counts[88].item())].cpu().numpy()
        gathered = sharding.gather_notes(local, rank * C * T_FRAMES)

    if rank == 0:
        tf_peak, hbm_peak, peak_src = measured_peaks()
        fl = stage_flops(C)
        per_stage = []
        for name, tot_ms, n in stages:
            avg = tot_ms / max(n, 1)
            ent = {"stage": name, "ms_per_launch": round(avg, 4), "launches": n}
            if name in fl and avg > 0:
                ent["tflops"] = round(fl[name] / (avg * 1e-3) / 1e12, 2)
            per_stage.append(ent)
        logmel_bytes = C * (N_SAMPLES * 4 + N_MELS * T_FRAMES * 4)
        per_stage.append({"stage": "frontend.logmel(3 launches)", "ms_per_launch": round(ms_logmel, 4), "launches": args.steps,
                          "algorithmic_GBps": round(logmel_bytes / (ms_logmel * 1e-3) / 1e9, 1),
                          "hbm_frac": round(logmel_bytes / (ms_logmel * 1e-3) / 1e9 / hbm_peak, 4)})
        per_stage.append({"stage": "post.sigmoid+notes(3 launches)", "ms_per_launch": round(ms_post, 4), "launches": args.steps})
        per_stage.sort(key=lambda e: -e["ms_per_launch"])
        gemm_like = [e for e in per_stage if "tflops" in e and not e["stage"].endswith(".rec")]
        top = gemm_like[0] if gemm_like else None
        roofline = None
        if top:
            st = top["stage"]
            kern = ("conv_halo_kernel (tcgen05/TMA halo-tile implicit-GEMM conv)" if st.startswith(("res", "freq", "c2"))
                    else "attention_tc_kernel (tcgen05 fused clamped-softmax attention)" if st == "attn.core"
                    else "tc_gemm_kernel (tcgen05/TMA persistent GEMM)")
            roofline = {"bound": "tensor", "kernel": f"{kern}, stage {st}",
                        "achieved": top["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": round(top["tflops"] / tf_peak, 4),
                        "traffic": ncu_traffic(st, C), "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                        "algorithmic_flops_per_launch": fl[st], "ms_per_launch": top["ms_per_launch"],
                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that launch, profiles/r1_traffic_chunks64.json",
                        "note": "peak = cuBLAS bf16 8192^3 sustained rate on this pool (MEASURED_PEAKS.json); frac > 1 means this "
                                "kernel runs faster than that GEMM does back to back"}
            burst = measured_burst()
            if burst:
                roofline["frac_of_burst_peak"] = round(top["tflops"] / burst, 4)
        total_flops = sum(stage_flops(C).values())
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(C, world),
                "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(ms_e2e / args.steps, 3), "sm_mhz": clocks_e2e["sm_mhz"]},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "model_tflops_whole_step": round(total_flops * world * args.steps / (ms / 1e3) / 1e12, 2),
                "stages": per_stage, "notes_last_step": n_notes,
                "gathered_notes": None if gathered is None else int(len(gathered))}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_reference_chunks_per_sec(1, threads)                       # warm-up
            v, secs = cpu_reference_chunks_per_sec(args.cpu_baseline_chunks, threads)
            line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_baseline_chunks} chunks, B=1 loop as main.py:258-266 ({secs:.1f} s), oracle port, "
                                              f"torch {torch.get_num_threads()} threads"}
        if world == 1 and args.gpu_eager_baseline > 0:
            n = min(args.gpu_eager_baseline, C)
            v, ms_eager = gpu_eager_chunks_per_sec(wav, n)
            line["gpu_eager_baseline"] = {"value": round(v, 2), "unit": UNIT, "batch_chunks": n, "ms_per_batch": round(ms_eager, 2),
                                          "kind": "oracle port, torch fp32 eager (cuDNN/cuBLAS, TF32 off), forward + sigmoid + threshold only"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: 30-s chunks/sec, audio -> piano-roll -> MIDI notes, CNNRNNModelLarge (BASELINE.json configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default): the synthetic 2-hour recording of BASELINE configs[3] -- 240 x 30-s chunks, SURVEY.md 8(d) chord
recipe k = 0..239 stored as 16-bit PCM -- through CNNRNNModelLarge (89 M: n_mels 320, hidden 512, 3 layers) to one MIDI
note list.  STRONG scaling: rank r owns the contiguous block shard_range(240, r, N) (30 chunks per GPU at N = 8), runs
log-mel -> forward -> sigmoid -> threshold -> note grouping on it in batches of <= 64 chunks, and ONE step ends when
rank 0 holds the stitched note list of the whole recording on the host (per-rank lists all-gathered over NCCL, seam
notes merged: sharding.gather_notes_device).  `--chunks C` switches to weak scaling (C chunks per GPU per step).

  value : chunks/s of the whole job, inputs (float waveforms) resident in HBM, CUDA events, max over ranks
  e2e   : the same job through the public streaming API (pipeline.StreamingTranscriber) with HOST buffers: pinned 16-bit
          PCM -> H2D -> on-device conversion -> path -> D2H of the bit-packed piano-rolls and the note lists, every
          step, inside the timed region (copies of neighbouring batches overlap the compute: 3 streams, 2 slots)
  roofline     : the dominant tensor kernel of the step (per-stage CUDA events in a second pass of K steps)
  cpu_baseline : the oracle port (the reference's algorithm on the host cores), bounded sample, set-up outside the timer
  configs      : the other BASELINE configs measured beside it (N = 1): [0] CNNRNNModel 36 M one chunk (GPU + CPU),
                 [1] log-mel alone on 64 chunks, [2] Large forward at batch 16, [4] the 50 x 100 threshold-count sweep;
                 plus the precise (split-bf16) mode and the reference modules in torch eager on the same GPU.

--impl reference times the CPU path alone (the reference is pure Python/PyTorch + librosa and cannot be pip-installed
offline -- DESIGN.md -- so the arm runs the oracle port, the one other place that may execute oracle/).
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
GAIN = 3 ** -0.5          # default-init-scale weights (torch's own init variance): realistic logit statistics


# ------------------------------------------------------------------ algorithmic work per stage
def stage_flops(B: int, hidden: int = HIDDEN, large: bool = True) -> dict:
    """Algorithmic (unpadded) FLOPs of the GEMM-class stages for B chunks."""
    T = T_FRAMES
    F1, F2, F3 = 160, 80, 40
    H, Hl, D = hidden, hidden // 2, 3 * hidden
    conv = lambda F, co, ci, k: 2.0 * F * T * co * ci * k
    if not large:
        f = {"c2": conv(F1, 64, 32, 9), "rnn0.gemm": 2.0 * T * 8 * H * 64 * F2, "rnn1.gemm": 2.0 * T * 8 * H * 2 * H,
             "rnn2.gemm": 2.0 * T * 8 * H * 2 * H, "rnn0.rec": 2.0 * T * 8 * H * H, "rnn1.rec": 2.0 * T * 8 * H * H,
             "rnn2.rec": 2.0 * T * 8 * H * H, "heads": 2.0 * T * 88 * 2 * H}
        return {k: v * B for k, v in f.items()}
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
    """DRAM bytes of one launch of `stage` from the newest committed ncu --set full capture at the same chunk count."""
    for name in ("r2_traffic.json", "r1_traffic_chunks64.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            if d.get("chunks") == chunks and stage in d["dram_bytes_per_launch"]:
                return d["dram_bytes_per_launch"][stage], f"profiles/{name}"
        except Exception:
            pass
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops"), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 1400.0, None, 6650.0, "fallback (B200_PROFILING.md)"


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
class CpuReference:
    """The reference's own B = 1 loop (main.py:258-275) restated in oracle/: log-mel -> fp32 forward -> sigmoid ->
    threshold -> grouping, on the host cores.  Everything main.py does ONCE per run (model construction, main.py:41-54;
    here also the filterbank, which librosa caches) happens in the constructor, outside any timer."""

    def __init__(self, model_type: str, threads: int, ks):
        from music_transcription_b200 import synth
        from oracle import frontend as ofe
        torch.set_num_threads(threads)
        self.model_type, self.threads = model_type, threads
        self.sd = synth.synth_state_dict(model_type, N_MELS, HIDDEN, LAYERS, seed=1, gain=GAIN)
        self.fb = ofe.mel_filterbank(n_mels=N_MELS)
        self.waves = [synth.piano_chord(int(k)) for k in ks]
        self.cache = {}

    def run(self, n_chunks: int) -> float:
        """Seconds for n_chunks chunks (cycling through the prepared waveforms), B = 1 loop + grouping."""
        from oracle import frontend as ofe, model as omodel, notes as onotes
        torch.set_num_threads(self.threads)
        t0 = time.perf_counter()
        rolls = []
        for i in range(n_chunks):
            y = self.waves[i % len(self.waves)]
            mel = torch.from_numpy(ofe.logmel(y, n_mels=N_MELS, fb=self.fb))[None, None]
            if self.model_type == "cnn_rnn_large":
                logits = omodel.large_forward(self.sd, mel, HIDDEN, LAYERS, lstm_cache=self.cache)
            else:
                logits = omodel.small_forward(self.sd, mel, HIDDEN, LAYERS, lstm_cache=self.cache)
            rolls.append(onotes.threshold_roll(torch.sigmoid(logits)[0].numpy(), 0.5))
        onotes.group_notes(onotes.combine_piano_rolls(rolls))
        return time.perf_counter() - t0


def gpu_eager_chunks_per_sec(mel: torch.Tensor, sd_cpu, reps: int = 3):
    """SURVEY.md 8(d) 'GPU baseline beside it': the reference's modules as plain torch fp32 eager on this GPU (oracle
    port: F.conv2d / nn.LSTM / matmul -> cuDNN and cuBLAS kernels) with torch's DEFAULT TF32 flags -- the reference
    sets none --, fed the log-mel of our frontend, batched like our step.  A reported baseline, never on the product path."""
    from oracle import model as omodel
    dev = mel.device
    sd = {k: v.to(dev) for k, v in sd_cpu.items()}
    cache, best = {}, None
    for _ in range(reps + 1):                              # first pass = warm-up (cuDNN autotune, allocator, LSTM build)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        logits = omodel.large_forward(sd, mel, HIDDEN, LAYERS, lstm_cache=cache)
        roll = (torch.sigmoid(logits) > 0.5).float()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    del roll, sd, cache
    return mel.shape[0] / (best / 1e3), best


class HangGuard(threading.Thread):
    """Safety net for the two-lane mode.  Two forwards in flight is stable (soak: 8500 batches, profiles/r2_two_lanes.md), but
    THREE lanes, or two lanes with capped tensor grids, stop making progress on this hardware -- an unexplained
    scheduling cliff one step away.  While armed, the guard watches a progress stamp; if the GPU phases make no progress for
    `limit` seconds it re-executes this benchmark with --lanes 1 (exec tears the CUDA context down) instead of hanging."""

    def __init__(self, limit=60.0):
        super().__init__(daemon=True)
        self.limit, self.armed, self.stamp = limit, False, time.time()

    def tick(self):
        self.stamp = time.time()

    def arm(self, on=True):
        self.stamp, self.armed = time.time(), on

    def run(self):
        while True:
            time.sleep(2.0)
            if self.armed and time.time() - self.stamp > self.limit:
                sys.stderr.write(f"[bench] no GPU progress for {self.limit:.0f} s with --lanes 2: re-executing with --lanes 1\n")
                sys.stderr.flush()
                argv, skip = [], False
                for a in sys.argv:
                    if skip or a == "--lanes":
                        skip = a == "--lanes"             # drop the flag and its value
                        continue
                    if not a.startswith("--lanes="):
                        argv.append(a)
                os.execv(sys.executable, [sys.executable] + argv + ["--lanes", "1", "--fell-back"])


def workload_config(args, world, n_local, batches):
    if args.chunks:
        what = (f"weak scaling: {args.chunks} synthetic 30-s chunks per GPU per step (BASELINE configs[3]'s model and path, "
                "fixed per-GPU work)")
    else:
        what = (f"BASELINE configs[3]: synthetic 2-hour recording = {args.recording} x 30-s chunks (SURVEY 8d chord recipe "
                f"k = 0..{args.recording - 1}, 16-bit PCM), CNNRNNModelLarge (89M, n_mels 320, hidden 512, 3 layers) audio -> "
                "piano-roll -> MIDI note list, chunk-sharded across the GPUs (strong scaling); a step ends when rank 0 holds "
                "the stitched note list of the whole recording")
    return {"workload": what, "recording_chunks": args.chunks * world if args.chunks else args.recording,
            "chunks_per_gpu": n_local, "batches_per_gpu": [b - a for a, b in batches], "samples_per_chunk": N_SAMPLES,
            "frames": T_FRAMES, "threshold": 0.5, "precision": args.precision, "lanes": args.lanes,
            "parallelism": f"chunk-sharded x{world}",
            "l2": "each batch streams ~190 MB of activations per chunk (>> 126 MB L2) and every step re-reads its 1.9 MB/chunk "
                  "inputs after them; no explicit flush"}


def run_reference(args, rank, world):
    """Reference arm: the oracle port on the host cores, all threads; each step a bounded sample (--ref-chunks chunks of the
    recording, B = 1 loop).  Set-up (weights, filterbank, waveforms, module construction) is outside the timed steps."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_chunks
    ref = CpuReference("cnn_rnn_large", threads, range(n))
    ref.run(1)                                             # builds the cached LSTM modules (main.py builds its model once)
    for _ in range(args.warmup):
        ref.run(1)
    secs = [ref.run(n) for _ in range(args.steps)]
    dt = sum(secs)
    v = n * args.steps / dt
    n_local = args.chunks or args.recording
    from music_transcription_b200 import sharding
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.chunks else "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, max(args.gpus, 1), n_local, sharding.batch_ranges(n_local, args.batch)),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} chunk(s)/step x {args.steps} steps of the recording, B=1 loop as main.py:258-266, oracle "
                                       "port (numpy log-mel + fp32 torch forward + numpy grouping), all host threads; set-up "
                                       "(weights, filterbank, waveforms, nn.LSTM construction) outside the timed steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--recording", type=int, default=240, help="chunks of the synthetic recording (strong scaling; configs[3])")
    ap.add_argument("--chunks", type=int, default=0, help="weak scaling instead: this many chunks per GPU per step")
    ap.add_argument("--batch", type=int, default=64, help="largest batch of chunks one pass of the kernels takes")
    ap.add_argument("--precision", default="fast", choices=["fast", "precise"])
    ap.add_argument("--lanes", type=int, default=2, choices=[1, 2], help="batches in flight per GPU (2: recurrences of one overlap the tensor kernels of the other)")
    ap.add_argument("--fell-back", action="store_true", help=argparse.SUPPRESS)      # set by HangGuard's re-exec
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-chunks", type=int, default=4, help="chunks per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-chunks", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the side measurements (other BASELINE configs, precise mode, eager)")
    ap.add_argument("--no-verify", action="store_true", help="skip the gathered == single-rank check (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
        args.gpus = world

    import torch.distributed as dist
    from music_transcription_b200 import _lib, evaluate, pipeline, sharding, synth
    from music_transcription_b200.transcription_model import TranscriptionModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().amt_device_check())
    L = _lib.lib()
    stream = _lib.stream_ptr(dev)
    T = T_FRAMES

    # ---- the recording and this rank's block of it
    R = args.chunks * world if args.chunks else args.recording
    lo, hi = sharding.shard_range(R, rank, world)
    n_local = hi - lo
    batches = sharding.batch_ranges(n_local, args.batch)
    B0 = batches[0][1] - batches[0][0]
    # 16-bit PCM of the chord recipe, as a WAVE file of the recording stores it; the float waveform of the same samples
    # (what librosa.load hands main.py:76) is the device-resident input of `value`
    pcm_host = synth.to_pcm16(synth.piano_chord_batch_fast(range(lo, hi))).pin_memory()
    wav = (pcm_host.to(dev).float() / 32768.0).contiguous()

    sd_cpu = synth.synth_state_dict("cnn_rnn_large", N_MELS, HIDDEN, LAYERS, seed=1, gain=GAIN)
    model = TranscriptionModel("cnn_rnn_large", n_mels=N_MELS, hidden_size=HIDDEN, num_layers=LAYERS, dropout=0.2, device=dev,
                               precision=args.precision)
    model.load_state_dict(sd_cpu)
    model.eval()
    fe = pipeline.Frontend.get(device=dev)
    cap = 88 * ((n_local * T + 1) // 2)
    probs = torch.empty(n_local, 88, T, device=dev)
    rolls = torch.empty(n_local, 88, (T + 31) // 32, dtype=torch.int32, device=dev)
    notes = torch.empty(cap, 3, dtype=torch.int32, device=dev)
    counts = torch.empty(89, dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * 88 * n_local, dtype=torch.int32, device=dev)
    result = {}
    W32 = (T + 31) // 32
    # two LANES: consecutive batches run on two streams, so the latency-bound recurrences of one batch (50-100 idle SMs)
    # overlap the tensor kernels of the other; results are bitwise those of a single stream (tests)
    lane_streams = pipeline.lane_streams(dev) if args.lanes == 2 else [None]
    rolls2 = [torch.empty(n_local, 88, W32, dtype=torch.int32, device=dev) for _ in range(2)]
    gather_dev = sharding.AsyncRollGather(n_local, R, T, dev)
    counter = {"batch": 0}

    def launch_step(k):
        """All of this rank's batches of recording k -> packed rolls (sigmoid + strict float32 '>' of main.py:153-156 straight
        into bits, 10.6 KB per chunk), then the one exchange of the path, asynchronously: all-gather the packed rolls over
        NCCL / NVLink, ONE grouping pass over the gathered roll on the GPU (seams between chunks, batches and ranks merge in
        the kernel), note list -> pinned host memory on every rank."""
        out, evs = rolls2[k & 1], []
        for a, b in batches:
            lane = lane_streams[counter["batch"] % len(lane_streams)]
            counter["batch"] += 1
            cur = lane if lane is not None else torch.cuda.current_stream(dev)
            with torch.cuda.stream(cur):
                logits = model(fe.logmel(wav[a:b], defer_floor=True))
                _lib.check(L.amt_pack_roll_u32(_lib.ptr(logits), (b - a) * 88, T, 0.5, 1, _lib.ptr(out[a:b]), cur.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(cur)
                evs.append(ev)
        return gather_dev.submit(out, after=evs)

    def run_device(steps):
        """K recordings back to back; recording k+1 is launched before the note list of recording k is collected (the host
        never idles the GPU), and every recording's list reaches the host inside the timed region."""
        pending = None
        for k in range(steps):
            t = launch_step(k)
            if pending is not None:
                result["notes"] = gather_dev.result(pending)
                guard.tick()
            pending = t
        result["notes"] = gather_dev.result(pending)
        guard.tick()

    def step_sequential():
        """The same work on ONE stream with a synchronous exchange -- used for the per-stage event pass (clean kernel times)."""
        for a, b in batches:
            logits = model(fe.logmel(wav[a:b], defer_floor=True))
            _lib.check(L.amt_pack_roll_u32(_lib.ptr(logits), (b - a) * 88, T, 0.5, 1, _lib.ptr(rolls[a:b]), stream))
        result["notes_seq"] = sharding.gather_rolls_notes(rolls, T, R)

    # End to end through the public API with HOST buffers: pipeline.StreamingTranscriber copies batch i+1 in and batch i-1
    # out while batch i computes; every batch still moves its own audio in and its rolls + notes out.
    streamer = pipeline.StreamingTranscriber(model, B0, N_SAMPLES, 0.5, input_format="pcm16", roll_format="bits", lanes=args.lanes)
    e2e_stats = {"d2h": 0}

    rec_bits = torch.zeros(n_local, 88, (T + 31) // 32, dtype=torch.int32)        # this rank's packed rolls of one recording (host)
    gatherer = sharding.AsyncRollGather(n_local, R, T, dev)

    def run_e2e(steps):
        """Batch after batch through the streamer; when a recording's last batch has arrived on the host, its packed rolls go
        to the asynchronous roll exchange (upload 10.6 KB per chunk, all-gather, one grouping pass, note list back) on a side
        stream, and the PREVIOUS recording's note list is collected -- the host never waits behind queued batches."""
        feed = (pcm_host[a:b] for _ in range(steps) for a, b in batches)
        d2h, pending = 0, None
        for i, (roll, nts) in enumerate(streamer.run(feed)):
            a, b = batches[i % len(batches)]
            rec_bits[a:b].copy_(roll)
            guard.tick()
            d2h += roll.numel() * 4 + 89 * 4 + nts.size * 4                      # what the streamer downloaded for this batch
            if i % len(batches) == len(batches) - 1:                              # the recording's last batch on this rank
                ticket = gatherer.submit(rec_bits)
                if pending is not None:
                    result["e2e_notes"] = gatherer.result(pending)
                pending = ticket
        if pending is not None:
            result["e2e_notes"] = gatherer.result(pending)
            d2h += (result["e2e_notes"].size + 89) * 4 * steps                    # the recording's note list, once per step
        e2e_stats["d2h"] = d2h // max(steps, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, loop=True):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if loop:
            for _ in range(steps):
                fn()
        else:
            fn(steps)                        # fn pipelines the steps itself and returns with every result on the host
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    guard = HangGuard()
    if args.lanes == 2:
        guard.start()
        guard.arm()
    run_device(max(args.warmup, 1))
    torch.cuda.synchronize()
    guard.tick()

    # ---- end to end with host buffers (the headline; measured first, right after warm-up)
    run_e2e(1)
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
    e2e_value = R * args.steps / (ms_e2e / 1e3)
    io = torch.tensor([n_local * N_SAMPLES * 2 + rec_bits.numel() * 4, e2e_stats["d2h"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(io)                                                      # whole-job bytes per step (all ranks)
    h2d, d2h = int(io[0].item()), int(io[1].item())

    # ---- timed region (device-resident inputs): exactly K steps, clocks sampled, no per-stage events
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.amt_launch_count()
    ms = timed(run_device, args.steps, loop=False)
    launches = int(L.amt_launch_count() - launches0)
    clocks = sampler.stop()
    value = R * args.steps / (ms / 1e3)
    n_notes = int(len(result["notes"]))
    same = bool(np.array_equal(result["notes"], result["e2e_notes"]))           # the two paths saw the same samples

    # ---- the same K steps again with CUDA events around every kernel launch: the per-stage breakdown
    guard.arm(False)                                   # everything below runs on one stream
    step_sequential()
    ms_seq = timed(step_sequential, args.steps)
    seq_same = bool(np.array_equal(result["notes_seq"], result["notes"]))
    model.profile(True)
    timed(step_sequential, args.steps)
    stages = model.profile_read()
    model.profile(False)

    # ---- the stages outside amt_model_forward, timed alone (same stream, CUDA events)
    a0, b0 = batches[0]
    mel_keep = fe.logmel(wav[a0:b0])
    logits_keep = model(mel_keep)
    bits = torch.empty(B0, 88, (T + 31) // 32, dtype=torch.int32, device=dev)

    def post():
        _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits_keep), logits_keep.numel(), 0.5, _lib.ptr(probs[a0:b0]), 0, stream))
        _lib.check(L.amt_pack_roll_u32(_lib.ptr(probs[a0:b0]), B0 * 88, T, 0.5, 0, _lib.ptr(bits), stream))
        pipeline.extract_notes_async(probs[a0:b0], 0.5, notes, counts, scratch)
    ms_logmel = timed(lambda: fe.logmel(wav[a0:b0], defer_floor=True), args.steps) / args.steps      # as the step runs it
    ms_post = timed(post, args.steps) / args.steps
    del mel_keep, logits_keep

    # ---- N > 1: the gathered list must equal what ONE rank computes for the whole recording (outside the timer)
    verified = None
    if world > 1 and not args.no_verify:
        if rank == 0:
            full = synth.to_pcm16(synth.piano_chord_batch_fast(range(R))).to(dev).float() / 32768.0
            want, _ = pipeline.transcribe_chunks(model, full, threshold=0.5, batch=args.batch)
            verified = bool(np.array_equal(want, result["notes"]))
            del full
        dist.barrier()

    # ---- BASELINE configs[4]: 50 pieces x 100 thresholds, pieces sharded over the ranks, counts all-gathered
    sweep = None
    if not args.no_configs:
        n_pieces, thr = 50, np.linspace(0.01, 0.99, 100)
        plo, phi = sharding.shard_range(n_pieces, rank, world)
        lens_all = np.array([937, 938, 469] * 17, dtype=np.int32)[:n_pieces]
        P = torch.stack([torch.from_numpy(synth.planted_probs(88, T, thr, seed=i)) for i in range(plo, phi)]).to(dev)
        Y = torch.stack([torch.from_numpy(synth.bernoulli_roll(88, T, 0.05, seed=i)) for i in range(plo, phi)]).to(dev)
        lens = torch.from_numpy(lens_all[plo:phi]).to(dev)
        sw = {}

        def sweep_step():
            sw["counts"] = sharding.gather_counts(evaluate.f1_counts_device(P, Y, lens, thr), n_pieces)
        sweep_step()
        ms_sweep = timed(sweep_step, args.steps) / args.steps
        tot = sw["counts"]
        cells = int((lens_all.astype(np.int64) * 88).sum())
        sweep = {"config": "BASELINE configs[4]: framewise TP/FP/FN, 50 pieces x 100 thresholds, pieces sharded, counts all-gathered",
                 "ms": round(ms_sweep, 4), "piece_threshold_evals_per_s": round(n_pieces * 100 / (ms_sweep * 1e-3), 1),
                 "algorithmic_bytes": cells * 8, "checksum_tp_fp_fn": [int(x) for x in tot.sum(axis=(0, 1))],
                 "tp_plus_fn_equals_positives": bool((tot[:, :, 0] + tot[:, :, 2] == tot[:, :1, 0] + tot[:, :1, 2]).all())}

    if rank == 0:
        tf_peak, tf_burst, hbm_peak, peak_src = measured_peaks()
        fl = stage_flops(B0)
        per_stage = []
        for name, tot_ms, n in stages:
            avg = tot_ms / max(n, 1)
            ent = {"stage": name, "ms_per_launch": round(avg, 4), "launches": n}
            if name in fl and avg > 0:
                ent["tflops"] = round(fl[name] / (avg * 1e-3) / 1e12, 2)
            per_stage.append(ent)
        logmel_bytes = B0 * (N_SAMPLES * 4 + N_MELS * T * 4)
        per_stage.append({"stage": "frontend.logmel", "ms_per_launch": round(ms_logmel, 4), "launches": args.steps,
                          "algorithmic_GBps": round(logmel_bytes / (ms_logmel * 1e-3) / 1e9, 1),
                          "hbm_frac": round(logmel_bytes / (ms_logmel * 1e-3) / 1e9 / hbm_peak, 4)})
        per_stage.append({"stage": "post.sigmoid+pack+notes", "ms_per_launch": round(ms_post, 4), "launches": args.steps})
        per_stage.sort(key=lambda e: -e["ms_per_launch"])
        gemm_like = [e for e in per_stage if "tflops" in e and not e["stage"].endswith(".rec")]
        top = gemm_like[0] if gemm_like else None
        roofline = None
        if top:
            st = top["stage"]
            kern = ("conv_halo_kernel (tcgen05/TMA halo-tile implicit-GEMM conv)" if st.startswith(("res", "freq", "c2"))
                    else "attention_tc_kernel (tcgen05 fused clamped-softmax attention)" if st == "attn.core"
                    else "tc_gemm_kernel (tcgen05/TMA persistent GEMM)")
            traffic, traffic_src = ncu_traffic(st, B0)
            roofline = {"bound": "tensor", "kernel": f"{kern}, stage {st}, {B0} chunks per launch",
                        "achieved": top["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": round(top["tflops"] / tf_peak, 4),
                        "traffic": traffic, "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                        "algorithmic_flops_per_launch": fl[st], "ms_per_launch": top["ms_per_launch"],
                        "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum of that launch, {traffic_src}",
                        "note": "peak = cuBLAS bf16 8192^3 sustained rate on this pool; frac > 1 means this kernel runs faster "
                                "than that GEMM does back to back (it is measured at a power-capped clock)"}
            if tf_burst:
                roofline["frac_of_burst_peak"] = round(top["tflops"] / tf_burst, 4)
        total_flops = sum(stage_flops(1).values()) * R
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak" if args.chunks else "strong", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "fast" else "bf16x3 (split-bf16 operands, fp32 accumulate)", "data": "synthetic",
                "config": workload_config(args, world, n_local, batches),
                "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(ms_e2e / args.steps, 3), "sm_mhz": clocks_e2e["sm_mhz"],
                        "api": "pipeline.StreamingTranscriber(input_format='pcm16', roll_format='bits') per batch + "
                               "sharding.AsyncRollGather per recording (packed rolls all-gathered, one grouping pass, note list to the "
                               "host of every rank, one recording behind the compute); bytes are whole-job totals per step (all ranks)",
                        "notes_equal_device_path": same},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "model_tflops_whole_step": round(total_flops * args.steps / (ms / 1e3) / 1e12, 2),
                "stages": per_stage, "notes_per_recording": n_notes, "gathered_equals_single_rank": verified,
                "fell_back_to_one_lane": bool(args.fell_back),
                "single_stream": {"value": round(R * args.steps / (ms_seq / 1e3), 3), "ms_per_step": round(ms_seq / args.steps, 3),
                                  "what": "the same steps on ONE stream with a synchronous exchange (lanes = 1, no step pipelining); the "
                                          "per-stage times and the roofline kernel's TFLOP/s come from this form", "notes_equal": seq_same}}
        side = {}
        if sweep:
            side["configs[4]_threshold_sweep"] = sweep
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ref = CpuReference("cnn_rnn_large", threads, range(min(args.cpu_baseline_chunks, 8)))
            ref.run(1)                                                          # warm-up + module construction
            secs = ref.run(args.cpu_baseline_chunks)
            line["cpu_baseline"] = {"value": round(args.cpu_baseline_chunks / secs, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_baseline_chunks} chunks of the recording, B=1 loop as main.py:258-266 ({secs:.1f} s), "
                                              f"oracle port, torch {torch.get_num_threads()} threads, set-up outside the timer"}
            del ref
        if world == 1 and not args.no_configs:
            side.update(side_measurements(args, dev, model, fe, wav, sd_cpu, hbm_peak, tf_peak))
        if side:
            line["configs"] = side
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def side_measurements(args, dev, model, fe, wav, sd_cpu, hbm_peak, tf_peak):
    """The other BASELINE configs on this GPU (N = 1 only), each timed with CUDA events over `steps` repetitions."""
    from music_transcription_b200 import _lib, pipeline, synth
    from music_transcription_b200.transcription_model import TranscriptionModel
    L = _lib.lib()
    stream = _lib.stream_ptr(dev)
    T = T_FRAMES
    out = {}

    def timed(fn, steps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # configs[1]: log-mel frontend alone, batch of 64 chunks
    w64 = wav[:64] if wav.shape[0] >= 64 else wav[:1].expand(64, -1).contiguous()
    ms = timed(lambda: fe.logmel(w64), args.steps)
    nbytes = 64 * (N_SAMPLES * 4 + N_MELS * T * 4)
    out["configs[1]_logmel_64_chunks"] = {"ms": round(ms, 4), "chunks_per_s": round(64 / (ms * 1e-3), 1), "algorithmic_bytes": nbytes,
                                          "GBps": round(nbytes / (ms * 1e-3) / 1e9, 1), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / hbm_peak, 4)}

    # configs[2]: CNNRNNModelLarge forward (log-mel -> three heads), batch 16
    mel16 = fe.logmel(wav[:16]).clone()
    ms = timed(lambda: model(mel16, return_all_heads=True), args.steps)
    fl = sum(stage_flops(16).values())
    out["configs[2]_large_forward_batch16"] = {"ms": round(ms, 4), "chunks_per_s": round(16 / (ms * 1e-3), 1),
                                               "model_tflops": round(fl / (ms * 1e-3) / 1e12, 1),
                                               "frac_of_sustained_bf16_peak": round(fl / (ms * 1e-3) / 1e12 / tf_peak, 4)}

    # precise (split-bf16) mode: the same forward + the 60-chunk batch of the headline step
    if args.precision == "fast":
        pm = TranscriptionModel("cnn_rnn_large", n_mels=N_MELS, hidden_size=HIDDEN, num_layers=LAYERS, dropout=0.2, device=dev,
                                precision="precise")
        pm.load_state_dict(sd_cpu)
        nb = min(60, wav.shape[0])
        probs = torch.empty(nb, 88, T, device=dev)

        def precise_step():
            logits = pm(fe.logmel(wav[:nb], defer_floor=True))
            _lib.check(L.amt_sigmoid_threshold(_lib.ptr(logits), logits.numel(), 0.5, _lib.ptr(probs), 0, stream))
        ms = timed(precise_step, max(args.steps // 2, 2))
        dp = (torch.sigmoid(pm(mel16)) - torch.sigmoid(model(mel16))).abs().max().item()
        out["precise_mode"] = {"ms_per_batch": round(ms, 3), "batch_chunks": nb, "chunks_per_s": round(nb / (ms * 1e-3), 1),
                               "what": "precision='precise': split-bf16 operands, 3 MMA products per contraction; log-mel -> forward -> sigmoid",
                               "max_abs_prob_diff_vs_fast_mode": dp}
        del pm, probs

    # configs[0]: CNNRNNModel (36 M), one chunk, B = 1, main.py path: log-mel -> forward -> sigmoid -> threshold -> notes
    small = TranscriptionModel("cnn_rnn", n_mels=N_MELS, hidden_size=HIDDEN, num_layers=LAYERS, dropout=0.2, device=dev)
    small.load_state_dict(synth.synth_state_dict("cnn_rnn", N_MELS, HIDDEN, LAYERS, seed=1, gain=GAIN))
    w1 = wav[:1]

    def small_step():
        roll = small.predict(fe.logmel(w1), threshold=0.5)
        return pipeline.extract_notes(roll[0], threshold=0.0)
    ms = timed(small_step, args.steps)
    ent = {"gpu_ms_per_chunk": round(ms, 4), "gpu_chunks_per_s": round(1e3 / ms, 1),
           "what": "CNNRNNModel 36M random-init, one synthetic 30-s chunk, B=1: log-mel -> forward -> sigmoid -> >0.5 -> grouping, "
                   "notes on the host (includes the host sync of every call)"}
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        for th in (threads, 1):
            ref = CpuReference("cnn_rnn", th, [0])
            ref.run(1)
            secs = min(ref.run(1) for _ in range(2 if th == 1 else 3))
            ent[f"cpu_port_chunks_per_s_{th}_threads"] = round(1.0 / secs, 4)
        torch.set_num_threads(threads)
    out["configs[0]_small_model_one_chunk"] = ent
    del small

    # SURVEY 8(d) "GPU baseline beside it": the reference's modules in torch eager on this GPU, default TF32 flags
    v, ms = gpu_eager_chunks_per_sec(mel16, sd_cpu)
    out["gpu_eager_baseline"] = {"value": round(v, 2), "unit": UNIT, "batch_chunks": 16, "ms_per_batch": round(ms, 2),
                                 "tf32_flags": {"matmul": bool(torch.backends.cuda.matmul.allow_tf32), "cudnn": bool(torch.backends.cudnn.allow_tf32)},
                                 "kind": "oracle port (the reference's modules), torch fp32 eager with torch's default TF32 flags "
                                         "(cuDNN / cuBLAS), forward + sigmoid + threshold only, same log-mel input"}
    return out


if __name__ == "__main__":
    main()

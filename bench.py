#!/usr/bin/env python
"""bench.py -- per-read methylation site calls/sec of the mCaller hot path on B200 (BASELINE.json metric).

A step = one pass of the hot path (TSV scan -> window features -> MLP -> per-site histogram) over one batch of
synthetic eventalign text: BASELINE.json configs[1] (synthetic E. coli 4.6 Mb, 100k reads, -m GATC, NN model, -n 6)
per GPU.  `value` times the pass with the text already resident in HBM; `e2e` times the same metric through the
public host-buffer path (pinned host memory -> H2D -> kernels -> rows D2H).  With --gpus N (torchrun) every rank owns
its own 100k-read slice (weak scaling); the per-site histograms are all-reduced over NCCL and the slice-edge window is
handed to the next rank inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "models", "r95_twobase_model_NN_6_m6A.pkl")
METRIC = "per-read site calls/sec"
UNIT = "calls/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--reads", type=int, default=100000, help="reads per GPU (BASELINE configs[1]: 100k)")
    p.add_argument("--e2e-bytes", type=float, default=8e9, help="size of the pinned host sample streamed by the e2e leg")
    p.add_argument("--e2e-chunk", type=int, default=1 << 30)
    p.add_argument("--cpu-reads", type=int, default=4000, help="reads in the bounded CPU-baseline sample")
    p.add_argument("--skip", type=int, default=0, help="-s skip threshold")
    p.add_argument("--classifier", default="NN", choices=["NN", "RF"],
                   help="NN: the shipped r95 MLP pickle (BASELINE configs[1]); RF: a forest with the reference's -c RF hyper-parameters "
                        "(train_model.py:39-45) fitted on synthetic features (configs[3], tree-walk kernel)")
    p.add_argument("--contig-name", default="ecoli", help="name of the synthetic contig (column 1 of every line)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    return p.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs: NVML (a query takes ~0.1 ms, so even a
    100 ms region gets hundreds of samples) when the bindings load, else the nvidia-smi query of the profiling recipe."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.source = index, [], False, "nvidia-smi"
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")) and index < len(vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown, n.nvmlClocksEventReasonSwThermalSlowdown,
                n.nvmlClocksEventReasonSwPowerCap]
        self.rows.append([str(mhz), str(self.max_mhz)] + ["Active" if (r & b) else "Not Active" for b in bits])

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                if self.nvml is not None:
                    self.nvml = None                     # fall back to nvidia-smi for the rest of the run
                    self.source = "nvidia-smi"
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = [n for j, n in enumerate(self.NAMES) if any(len(r) > 2 + j and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def fit_reference_rf():
    """{'MG': rf, 'MH': rf} with the hyper-parameters of the reference's -c RF (train_model.py:39-45 minus the arguments
    current scikit-learn has dropped), fitted on 40 000 synthetic feature rows -- no RF pickle is shipped."""
    from sklearn.ensemble import RandomForestClassifier
    rng = np.random.RandomState(0)
    X = np.column_stack([rng.normal(0.0, 2.4, size=(40000, 6)), rng.uniform(3.0, 24.0, size=40000)])
    y = np.where(X[:, 2] - 0.6 * X[:, 3] + 0.3 * rng.normal(size=40000) > 0.5, "m6A", "A")
    rf = RandomForestClassifier(n_estimators=50, criterion="entropy", max_depth=10, max_features=4, min_samples_leaf=2,
                                min_samples_split=3, random_state=0, n_jobs=-1).fit(X, y)
    return {"MG": rf, "MH": rf}


def build_world(args, rank, world, n_generate=None):
    """Reference index, models, quality table and this rank's synthetic text in HBM."""
    import torch
    from mcaller_b200 import engine as eng_mod, models, read_qual, refmark, synth, synth_device
    from mcaller_b200.refindex import ReferenceIndex
    spec = synth.SynthSpec(seed=0, contigs=[(args.contig_name, 4600000)], n_reads=args.reads * world, len_min=1000, len_max=3000)
    genome = synth.genome(spec, 0)
    seqs = {args.contig_name: genome.tobytes().decode()}
    ref = ReferenceIndex(seqs, "A", motif="GATC", k=6)
    meth = {0: (synth.meth_sites(spec, 0, ref.site_fwd_bits[:4600000]), synth.meth_sites(spec, 0, ref.site_rev_bits[:4600000]))}
    gen = synth_device.DeviceSynth(spec, ref, meth)
    lo, hi = rank * args.reads, (rank + 1) * args.reads
    if n_generate is not None:
        hi = lo + min(n_generate, hi - lo)
    d_text, nbytes, offs = gen.generate(lo, hi - lo)
    torch.cuda.synchronize()
    keys, q = synth_device.quality_table_for(spec, lo, hi)
    qt = read_qual.build_quality_table(dict(zip(keys, q.tolist())))
    model = models.load_model_file(MODEL) if args.classifier == "NN" else fit_reference_rf()
    e0, e1, two = models.select_models(model, "A")
    dm = models.DeviceModels(e0, e1)
    engine = eng_mod.Engine(ref, models=dm, qual_table=qt, skip_thresh=args.skip, qual_thresh=0.0, two_models=two, histogram=True)
    return dict(spec=spec, ref=ref, gen=gen, d_text=d_text, nbytes=nbytes, offs=offs, engine=engine, seqs=seqs, lo=lo, hi=hi,
                quals=dict(zip(keys, q.tolist())), model=model)


def cpu_oracle_pass(text_bytes, read_offsets, seqs, quals, model, skip, threads):
    """The CPU restatement (oracle/mcaller_oracle.c) over a host sample, read-aligned slices on `threads` host threads.
    Returns (calls, seconds).  Slices are independent files here: each drops its own last open window (reference Q3)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.lib()
    n = len(read_offsets)
    threads = max(1, min(threads, n))
    bounds = [read_offsets[(n * t) // threads] for t in range(threads)] + [len(text_bytes)]
    mv = memoryview(text_bytes)

    # reference marking, quality table and model are marshalled once (the reference also loads them once per worker)
    prep = orc.prepare(seqs, quals, model=model, base="A", motif="GATC")

    def work(t):
        sl = bytes(mv[bounds[t]:bounds[t + 1]])
        r = orc.extract(sl, None, None, k=6, skip_thresh=skip, qual_thresh=0.0, cap=max(4096, len(sl) // 2000), count_only=True,
                        prepared=prep)
        return r["counters"]["observations"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        calls = sum(ex.map(work, range(threads)))
    return calls, time.perf_counter() - t0


_REAL_STDOUT = None


def emit_line(obj):
    """The one JSON line of the contract goes to the real stdout; everything else (NCCL banners, library prints) was
    redirected to stderr at start-up."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # fd 1 -> stderr for the rest of the run
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    import torch
    from mcaller_b200 import _lib, dist as mdist, engine as eng_mod, stream as stream_mod
    eng_mod.require_cuda()
    torch.cuda.set_device(local_rank)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W = build_world(args, rank, 1 if args.impl == "reference" else world,
                    n_generate=min(args.cpu_reads, args.reads) if args.impl == "reference" else None)
    engine, d_text, nbytes = W["engine"], W["d_text"], W["nbytes"]
    dev = engine.device
    host_cores = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        n_s = W["hi"] - W["lo"]
        end = nbytes
        sample = d_text[:end].cpu().numpy().tobytes()
        roffs = [int(x) for x in W["offs"].cpu().numpy()]
        orc_model = W["model"]
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_oracle_pass(sample, roffs, W["seqs"], W["quals"], orc_model, args.skip, host_cores)
        tot_calls, tot_t = 0, 0.0
        for _ in range(args.steps):
            c, t = cpu_oracle_pass(sample, roffs, W["seqs"], W["quals"], orc_model, args.skip, host_cores)
            tot_calls += c
            tot_t += t
        v = tot_calls / tot_t
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic E. coli 4.6 Mb, 100k reads, -m GATC, %s, -n 6, -s %d"
                                       % ("NN model (r95)" if args.classifier == "NN" else "RF model (50 trees, depth 10)", args.skip),
                           "sample": "%d reads (%.2f GB of eventalign TSV) per step" % (n_s, end / 1e9)},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_cores, "kind": "port",
                                 "sample": "C restatement of the reference (oracle/mcaller_oracle.c) on %d reads, %d host threads; the "
                                           "reference itself is pure Python and cannot travel to this box" % (n_s, host_cores)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit_line(line)
        return 0

    # ------------------------------------------------------------------------------------------------ our arm
    scan_ms = []

    def step(record_scan=False):
        engine.reset_histogram()
        engine.row_base = mdist.rank_row_base(rank)
        if record_scan:
            engine.scan_events = []
        res = engine.run_chunk(d_text, nbytes)
        st = engine.count_rows(res)
        calls = st["calls"]
        if use_dist:
            fk = int(engine.records(1)[0]["contig"]) if res.n_records else -1
            resolved, _ = mdist.exchange_boundaries(fk, st["pending"], dev)
            calls += resolved
            mdist.allreduce_histogram(engine.d_depth, engine.d_meth, engine.d_first)
        if record_scan and engine.scan_events:
            torch.cuda.synchronize()
            scan_ms.extend(a.elapsed_time(b) for a, b in engine.scan_events)
            engine.scan_events = None
        return calls, res

    for _ in range(max(args.warmup, 3)):
        calls, res = step()
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = engine.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    total_calls = 0
    for _ in range(args.steps):
        c, res = step(record_scan=True)
        total_calls += c
    ev1.record()
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = engine.launches - launches0
    clocks = sampler.summary()
    if use_dist:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
        c = torch.tensor([total_calls], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_calls = int(c[0])
    value = total_calls / (elapsed_ms / 1e3)

    # roofline of the dominant kernel (k_scan): algorithmic bytes = the text once + 32 B per record written
    peak, peak_src = measured_peak_hbm()
    alg_bytes = nbytes + 32 * res.n_records
    scan_avg_ms = float(np.mean(scan_ms)) if scan_ms else None
    achieved = alg_bytes / (scan_avg_ms / 1e3) / 1e9 if scan_avg_ms else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "kernel": "k_scan", "kernel_ms": scan_avg_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src}
    tr = os.path.join(ROOT, "profiles", "k_scan_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_at_bench_size")
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------ e2e leg
    e2e = None
    if not args.no_e2e:
        offs_all = W["offs"].cpu().numpy()
        n_e = int(np.searchsorted(offs_all, args.e2e_bytes, side="right"))
        n_e = max(1, min(n_e, len(offs_all)))
        end = int(offs_all[n_e]) if n_e < len(offs_all) else nbytes
        host = torch.empty(end, dtype=torch.uint8, pin_memory=True)
        host.copy_(d_text[:end])
        torch.cuda.synchronize()
        streamer = stream_mod.HostStreamer(engine, chunk_bytes=min(args.e2e_chunk, max(end, 1 << 20)))
        cuts = stream_mod.plan_chunks(offs_all[:n_e], end, streamer.chunk_bytes)
        engine.reset_histogram()
        sink = stream_mod.TextSink(W["ref"], 6, "A")             # rows rendered as .diffs text on host threads, one chunk behind
        streamer.run(host, cuts, sink=sink)                        # warm-up (buffer growth, pinned result buffers)
        streamer.h2d_bytes = streamer.d2h_bytes = 0
        sink.text_bytes = 0
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e_calls = 0
        e_steps = max(1, min(args.steps, 3))
        for _ in range(e_steps):
            engine.reset_histogram()
            tot = streamer.run(host, cuts, sink=sink)
            e_calls += tot["calls"]
            if use_dist:
                mdist.allreduce_histogram(engine.d_depth, engine.d_meth, engine.d_first)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if use_dist:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
            c = torch.tensor([e_calls], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            e_calls = int(c[0])
        e2e = {"value": e_calls / dt, "unit": UNIT, "h2d_bytes_per_step": streamer.h2d_bytes // e_steps,
               "d2h_bytes_per_step": streamer.d2h_bytes // e_steps, "diffs_text_bytes_per_step": sink.text_bytes // e_steps,
               "sample": "%d reads (%.2f GB TSV) per GPU streamed from pinned host memory in %d chunks, rows copied back and "
                         "rendered as .diffs text by the native writer" % (n_e, end / 1e9, len(cuts))}
        del host

    # ------------------------------------------------------------------------------------------------ CPU baseline
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = min(args.cpu_reads, args.reads)
        offs_all = W["offs"].cpu().numpy()
        end = int(offs_all[n_s]) if n_s < len(offs_all) else nbytes
        sample = d_text[:end].cpu().numpy().tobytes()
        roffs = [int(x) for x in offs_all[:n_s]]
        cpu_oracle_pass(sample, roffs, W["seqs"], W["quals"], W["model"], args.skip, host_cores)          # warm-up pass
        c_calls, c_t = cpu_oracle_pass(sample, roffs, W["seqs"], W["quals"], W["model"], args.skip, host_cores)
        cpu = {"value": c_calls / c_t, "unit": UNIT, "cores": host_cores, "kind": "port",
               "sample": "%d reads (%.2f GB TSV), C restatement of the reference, %d host threads, %.1f s" % (n_s, end / 1e9, host_cores, c_t)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic E. coli 4.6 Mb, %d reads per GPU (%.1f GB eventalign TSV), -m GATC, %s, -n 6, -s %d"
                                       % (args.reads, nbytes / 1e9, "NN model (r95)" if args.classifier == "NN" else
                                          "RF model (50 trees, depth 10, fitted on synthetic features)", args.skip),
                           "l2": "inputs (%.1f GB) larger than L2 (126 MB); no flush needed" % (nbytes / 1e9),
                           "parallelism": "reads sharded over %d GPU(s); histogram all-reduce + slice-edge hand-off" % world,
                           "calls_per_step": total_calls // args.steps, "lines_per_step_per_gpu": res.counters["lines"],
                           "records_per_step_per_gpu": res.n_records},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
        emit_line(line)
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

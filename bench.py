#!/usr/bin/env python
"""bench.py -- reads/s of the re-alignment hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1]): C. elegans-size synthetic reference (100 Mbp, 6 contigs, repeat
families, N runs; tools/mmsynth.c, seed 42) and 2x150 bp FR pairs drawn from updated regions (seed 44+rank),
mapped with the `-ax sr` preset.  One step = one mini-batch of pairs through the whole hot path
(sketch -> seed -> chain -> hits -> ksw2 extension -> MAPQ/pairing), i.e. everything between
worker_pipeline's step 0 (parse) and step 2 (format) of the reference (map.c:590-593).

  value : reads/s with the batch already resident in HBM (mm_b200_map_batch mode 2)
  e2e   : reads/s through the batch C-ABI with HOST buffers (mode 0: H2D of the reads, all stages,
          D2H of chains / DP results, malloc'd mm_reg1_t out)
  roofline : the dominant kernel by CUDA-event time inside the timed region
  cpu_baseline / --impl reference : the reference fork (oracle/_ref/minimap2_B, built from /root/reference
          by oracle/Makefile) on the box's own cores, -t nproc, mapping phase only, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "airlift_b200", "libmm2b200.so")
SYNTH = os.path.join(ROOT, "build", "mmsynth")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minimap2_B")
CLI_BIN = os.path.join(ROOT, "build", "minimap2-b200")
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "traffic_r01.json")   # dram bytes per launch from the committed ncu --set full captures
# integer-pipe roofline of K4: 148 SMs x 128 lanes x 1.965 GHz lane-ops/s over ~30 lane-ops per DP cell (one cell per lane-op)
K4_INT_ROOFLINE_GCUPS = 148 * 128 * 1.965 / 30.0

GENOME_BP = 100_000_000
N_CONTIGS = 6
PAIRS_PER_STEP = 500_000          # 1 M reads = 150 Mbases per step: larger than the 126 MB L2
METRIC = "remapped reads/sec at 1/2/4/8 B200; ksw2 GCUPS; seed-lookup HBM GB/s"


class IdxOpt(C.Structure):
    _fields_ = [("k", C.c_short), ("w", C.c_short), ("flag", C.c_short), ("bucket_bits", C.c_short), ("mini_batch_size", C.c_int),
                ("batch_size", C.c_uint64)]


class MapOptFull(C.Structure):  # mm_mapopt_t (include/minimap_b200.h == minimap.h:107-150)
    _fields_ = [("flag", C.c_int64), ("seed", C.c_int), ("sdust_thres", C.c_int), ("max_qlen", C.c_int), ("bw", C.c_int),
                ("max_gap", C.c_int), ("max_gap_ref", C.c_int), ("max_frag_len", C.c_int), ("max_chain_skip", C.c_int),
                ("max_chain_iter", C.c_int), ("min_cnt", C.c_int), ("min_chain_score", C.c_int), ("mask_level", C.c_float),
                ("pri_ratio", C.c_float), ("best_n", C.c_int), ("max_join_long", C.c_int), ("max_join_short", C.c_int),
                ("min_join_flank_sc", C.c_int), ("min_join_flank_ratio", C.c_float), ("a", C.c_int), ("b", C.c_int), ("q", C.c_int),
                ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int), ("sc_ambi", C.c_int), ("noncan", C.c_int), ("junc_bonus", C.c_int),
                ("zdrop", C.c_int), ("zdrop_inv", C.c_int), ("end_bonus", C.c_int), ("min_dp_max", C.c_int), ("min_ksw_len", C.c_int),
                ("anchor_ext_len", C.c_int), ("anchor_ext_shift", C.c_int), ("max_clip_ratio", C.c_float), ("pe_ori", C.c_int),
                ("pe_bonus", C.c_int), ("mid_occ_frac", C.c_float), ("min_mid_occ", C.c_int32), ("mid_occ", C.c_int32),
                ("max_occ", C.c_int32), ("mini_batch_size", C.c_int), ("max_sw_mat", C.c_int64), ("split_prefix", C.c_char_p)]


class Stats(C.Structure):  # mm_b200_stats_t
    _fields_ = [(n, C.c_double) for n in ("t_total", "t_upload", "t_seedchain", "t_seedchain_kernels", "t_hits", "t_align_host",
                                          "t_ksw_total", "t_ksw_kernel", "t_finish")] + \
               [(n, C.c_uint64) for n in ("n_frag", "n_reads", "n_bases", "n_minimizers", "n_anchors", "n_chain_iter", "n_dp_jobs",
                                          "n_dp_cells", "n_dp_rounds", "h2d_bytes", "d2h_bytes", "n_dp_jobs_fast", "n_dp_cells_fast")]


def load_lib():
    if not os.path.exists(LIB):
        raise SystemExit(f"{LIB} missing: run `make` (the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(LIB)
    L.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.POINTER(MapOptFull)]
    L.mm_idx_reader_open.restype = C.c_void_p
    L.mm_idx_reader_open.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.c_char_p]
    L.mm_idx_reader_read.restype = C.c_void_p
    L.mm_idx_reader_read.argtypes = [C.c_void_p, C.c_int]
    L.mm_idx_reader_close.argtypes = [C.c_void_p]
    L.mm_idx_destroy.argtypes = [C.c_void_p]
    L.mm_mapopt_update.argtypes = [C.POINTER(MapOptFull), C.c_void_p]
    L.mm_b200_set_devices.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.mm_b200_open_reads.restype = C.c_void_p
    L.mm_b200_open_reads.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.mm_b200_close_reads.argtypes = [C.c_void_p]
    L.mm_b200_read_batch.restype = C.c_void_p
    L.mm_b200_read_batch.argtypes = [C.c_void_p, C.POINTER(MapOptFull), C.c_int]
    L.mm_b200_batch_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.mm_b200_map_batch.argtypes = [C.c_void_p, C.POINTER(MapOptFull), C.c_int, C.c_void_p, C.c_int]
    L.mm_b200_reset_batch.argtypes = [C.c_void_p]
    L.mm_b200_free_batch.argtypes = [C.c_void_p]
    L.mm_b200_batch_digest.restype = C.c_uint64
    L.mm_b200_batch_digest.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.mm_b200_stats.argtypes = [C.POINTER(Stats), C.c_int]
    L.mm_b200_profile.argtypes = [C.c_void_p, C.c_int]
    L.mm_b200_set_serial.argtypes = [C.c_int]
    L.mm_b200_profile_fetch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_long)]
    L.mm_b200_idx_image.argtypes = [C.c_void_p, C.c_void_p]
    L.mm_b200_idx_alloc.restype = C.c_void_p
    L.mm_b200_idx_alloc.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.c_void_p, C.c_void_p]
    L.mm_b200_idx_finalize.argtypes = [C.c_void_p]
    L.mmg_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.mm_b200_ctx.restype = C.c_void_p
    L.mm_b200_ctx.argtypes = [C.c_void_p, C.c_int]
    L.mmg_stream.restype = C.c_void_p
    L.mmg_stream.argtypes = [C.c_void_p]
    L.mm_b200_launch_count.restype = C.c_long
    L.mm_b200_launch_count.argtypes = [C.c_void_p, C.c_int]
    return L


def ensure_tools():
    if not os.path.exists(SYNTH):
        subprocess.check_call(["make", "-C", ROOT, "tools"], stdout=subprocess.DEVNULL)


def workdir(need_bytes=0):
    """/dev/shm when it can hold the synthetic files of all ranks (what is already there from an earlier run counts), else the
    temp dir.  Every rank evaluates this before anything is written, so all ranks pick the same place."""
    import shutil
    for base in (["/dev/shm"] if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else []) + [tempfile.gettempdir()]:
        d = os.path.join(base, "airlift_b200_bench")
        have = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) if os.path.isdir(d) else 0
        if base != "/dev/shm" or shutil.disk_usage(base).free + have >= need_bytes * 1.1 + (1 << 30):
            os.makedirs(d, exist_ok=True)
            return d


def make_ref(d):
    fa = os.path.join(d, f"ref_{GENOME_BP}.fa")
    if not os.path.exists(fa):
        tmp = fa + f".tmp{os.getpid()}"
        subprocess.check_call([SYNTH, "ref", tmp, str(GENOME_BP), str(N_CONTIGS), "42"])
        os.replace(tmp, fa)
    return fa


def make_reads(d, fa, n_pairs, seed, tag, read_seed=0):
    """seed fixes the updated regions (a property of the reference pair); read_seed, when given, the reads drawn from them"""
    f1, f2 = os.path.join(d, f"{tag}_1.fq"), os.path.join(d, f"{tag}_2.fq")
    if not (os.path.exists(f1) and os.path.exists(f2)):
        t1, t2 = f1 + f".tmp{os.getpid()}", f2 + f".tmp{os.getpid()}"
        subprocess.check_call([SYNTH, "sr", fa, t1, t2, str(n_pairs), str(seed)] + (["0.02", str(read_seed)] if read_seed else []))
        os.replace(t1, f1)
        os.replace(t2, f2)
    return f1, f2


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.max_mhz, self._halt = gpu_index, [], set(), None, threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def ref_mapping_phase(cmd):
    """Run the reference CLI; return seconds between `loaded/built the index` and `Real time`."""
    p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    t_idx = t_end = None
    for line in p.stderr.splitlines():
        m = re.match(r"\[M::main::([0-9.]+)\*", line)
        if m and "loaded/built the index" in line:
            t_idx = float(m.group(1))
        m = re.search(r"Real time: ([0-9.]+) sec", line)
        if m:
            t_end = float(m.group(1))
    if p.returncode != 0 or t_idx is None or t_end is None:
        raise RuntimeError("reference run failed: " + p.stderr[-400:])
    return t_end - t_idx


def run_reference(args, d, fa):
    """--impl reference: the reference fork on the host cores, each step a bounded sample of the workload."""
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/minimap2_B is not built (make -C oracle ref needs /root/reference)"}))
        return
    cores = os.cpu_count() or 1
    n_pairs = 200_000
    f1, f2 = make_reads(d, fa, n_pairs, 44, f"sr_ref_{n_pairs}")
    cmd = [REF_BIN, "-ax", "sr", "-t", str(cores), fa, f1, f2]
    times = []
    for i in range(args.warmup + args.steps):
        t = ref_mapping_phase(cmd)
        if i >= args.warmup:
            times.append(t)
    per_step = sum(times) / len(times)
    v = 2 * n_pairs / per_step
    sample = f"{n_pairs} pairs (2x150) of the same workload per step, mapping phase only (index build excluded), -t {cores}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
        "data": "synthetic", "config": workload_config(n_pairs),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(pairs_per_step):
    return {"workload": f"C. elegans-size synthetic pair ({GENOME_BP // 10**6} Mbp, {N_CONTIGS} contigs), 2x150bp reads from updated regions, "
                        f"minimap2 -ax sr; {pairs_per_step} pairs per step", "preset": "sr", "k": 21, "w": 11,
            "pairs_per_step": pairs_per_step,
            "shards": "one job sharded by reads: every rank maps reads drawn (read seed 44 + rank) from the same updated regions (seed 44)",
            "l2_policy": "inputs larger than L2 (150 MB of reads + index per step), a different batch every step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the whole-CLI (parse + map + SAM) timing at N=1")
    ap.add_argument("--data-seed", type=int, default=0, help="seed of the synthetic reads (default 44 + rank)")
    ap.add_argument("--lanes", type=int, default=2, help="shards (streams) per GPU a batch is cut into")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    ensure_tools()
    # 2 x ~330 bytes of FASTQ per pair, warm-up + three timed passes, every rank its own files
    d = workdir(world * args.pairs * (args.warmup + 3 * args.steps) * 700 if args.impl == "b200" else 0)
    if args.impl == "reference":
        if rank == 0:
            fa = make_ref(d)
            run_reference(args, d, fa)
        return

    cli_result = None
    if rank == 0 and world == 1 and not args.no_cli and os.path.exists(CLI_BIN):
        # the whole drop-in binary (FASTQ parse + all stages + SAM on stdout), timed exactly like the reference arm; it runs
        # before this process touches the GPU, so that the binary sees the device a user's run would see
        fa = make_ref(d)
        n_pairs = 4_000_000
        c1, c2 = make_reads(d, fa, n_pairs, 45, f"sr_cli_{n_pairs}")
        try:
            t = ref_mapping_phase([CLI_BIN, "-ax", "sr", "-t", str(os.cpu_count() or 1), "-K", "150M", fa, c1, c2])
            cli_result = {"value": 2 * n_pairs / t, "unit": "reads/s", "sample": f"{n_pairs} pairs, build/minimap2-b200 -ax sr -t {os.cpu_count()} -K 150M, "
                          "mapping phase (after the index is built) incl. FASTQ parsing, first-batch buffer growth and SAM output"}
        except Exception as e:  # noqa
            cli_result = {"value": None, "unit": "reads/s", "sample": f"failed: {e}"}

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        fa = make_ref(d)
    if world > 1:
        dist.barrier()
    fa = make_ref(d)
    n_steps_total = args.warmup + 3 * args.steps
    # one job, sharded: every rank maps reads of the same updated regions (seed 44); rank r draws its own reads
    seed, read_seed = args.data_seed or 44, (0 if args.data_seed or rank == 0 else 44 + rank)
    f1, f2 = make_reads(d, fa, args.pairs * n_steps_total, seed, f"sr_s{seed}r{read_seed}_{args.pairs}x{n_steps_total}", read_seed)

    L = load_lib()
    dev = (C.c_int * 1)(local_rank)
    L.mm_b200_set_devices(1, dev)
    L.mm_b200_set_lanes(args.lanes)
    ipt, opt = IdxOpt(), MapOptFull()
    L.mm_set_opt(None, C.byref(ipt), C.byref(opt))
    L.mm_set_opt(b"sr", C.byref(ipt), C.byref(opt))
    opt.flag |= 0x004 | 0x008  # -a: MM_F_CIGAR | MM_F_OUT_SAM
    t0 = time.time()
    bcast_bytes = 0
    if world == 1 or rank == 0:
        rd = L.mm_idx_reader_open(fa.encode(), C.byref(ipt), None)
        mi = L.mm_idx_reader_read(rd, 3)
        L.mm_idx_reader_close(rd)
        L.mm_mapopt_update(C.byref(opt), mi)
    if world > 1:
        # the index is built once (rank 0) and replicated: one NCCL broadcast per buffer over NVLink/NVSwitch;
        # nothing else is ever exchanged between ranks
        from airlift_b200 import dist as D
        img = D.IdxImage()
        if rank == 0:
            assert L.mm_b200_idx_image(mi, C.byref(img)) == 0
        shape = D.broadcast_object((img.shape_tuple(), opt.mid_occ) if rank == 0 else None)
        if rank != 0:
            shp = D.IdxImage.from_shape(shape[0])
            img = D.IdxImage()
            mi = L.mm_b200_idx_alloc(fa.encode(), C.byref(ipt), C.byref(shp), C.byref(img))
            opt.mid_occ = shape[1]
        bcast_bytes = D.broadcast_buffers(L, img, torch.device("cuda", local_rank), src=0)
        if rank != 0:
            L.mm_b200_idx_finalize(mi)
        torch.cuda.synchronize()
    t_index = time.time() - t0
    n_threads = args.threads or max(1, (os.cpu_count() or 1) // world)
    ctx = L.mm_b200_ctx(mi, 0)
    stream = torch.cuda.ExternalStream(L.mmg_stream(ctx), device=torch.device("cuda", local_rank))

    fns = (C.c_char_p * 2)(f1.encode(), f2.encode())
    reader = L.mm_b200_open_reads(2, fns)
    batch_bases = args.pairs * 300

    def next_batch():
        b = L.mm_b200_read_batch(reader, C.byref(opt), batch_bases)
        if not b:
            raise SystemExit("ran out of synthetic reads")
        return b

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (full path, host buffers)
    for _ in range(args.warmup):
        b = next_batch()
        if L.mm_b200_map_batch(mi, C.byref(opt), n_threads, b, 0) != 0:
            raise SystemExit("mapping failed")
        L.mm_b200_free_batch(b)

    def timed(mode_resident):
        """K steps; returns (seconds max over ranks, reads on this rank, stats, kernel profile, launches, clocks)."""
        batches = [next_batch() for _ in range(args.steps)]
        n_reads = 0
        for b in batches:
            ns = C.c_int(0)
            L.mm_b200_batch_info(b, C.byref(ns), None, None)
            n_reads += ns.value
        L.mm_b200_stats(None, 1)
        L.mm_b200_launch_count(mi, 1)
        L.mm_b200_profile(mi, 1)
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev_ms = 0.0
        sync_all()
        t_wall0 = time.perf_counter()
        if not mode_resident:
            ev0.record(stream)
            for b in batches:
                if L.mm_b200_map_batch(mi, C.byref(opt), n_threads, b, 0) != 0:
                    raise SystemExit("mapping failed")
            ev1.record(stream)
            sync_all()
            dev_ms = ev0.elapsed_time(ev1)
        else:
            for b in batches:  # stage the reads in HBM outside the timed region, then time the resident pass
                if L.mm_b200_map_batch(mi, C.byref(opt), n_threads, b, 1) != 0:
                    raise SystemExit("upload failed")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if L.mm_b200_map_batch(mi, C.byref(opt), n_threads, b, 2) != 0:
                    raise SystemExit("mapping failed")
                e1.record(stream)
                torch.cuda.synchronize()
                dev_ms += e0.elapsed_time(e1)
            sync_all()
        wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
        st = Stats()
        L.mm_b200_stats(C.byref(st), 0)
        names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_long * 64)()
        nk = L.mm_b200_profile_fetch(mi, 64, names, ms, ln)
        prof = {}
        for i in range(nk):  # template instantiations (k_ksw<16>, k_ksw<32>) count as one kernel
            nm = names[i].decode().split("<")[0]
            prof[nm] = (prof.get(nm, (0.0, 0))[0] + ms[i], prof.get(nm, (0.0, 0))[1] + ln[i])
        L.mm_b200_profile(mi, 0)
        launches = L.mm_b200_launch_count(mi, 0)
        digest = 0
        for b in batches:
            digest ^= L.mm_b200_batch_digest(b, None)
            L.mm_b200_free_batch(b)
        secs = dev_ms * 1e-3
        if world > 1:
            t = torch.tensor([secs, float(n_reads)], device="cuda", dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            per_rank = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(per_rank, t)
            timed.per_rank_ms = [float(x[0]) * 1e3 / args.steps for x in per_rank]
            secs, total_reads = float(tmax[0]), float(tsum[1])
        else:
            total_reads = float(n_reads)
            timed.per_rank_ms = [secs * 1e3 / args.steps]
        return secs, total_reads, st, prof, launches, clocks, wall

    secs_e2e, reads_e2e, st_e2e, prof_e2e, launches_e2e, clocks_e2e, wall_e2e = timed(False)
    secs_res, reads_res, st_res, prof_val, launches_res, clocks_res, wall_res = timed(True)
    per_rank_ms = timed.per_rank_ms
    # Kernel profile: a third pass in which the shards of a GPU take turns on the device.  In the timed passes the two shards'
    # kernels overlap on purpose, which stretches every per-kernel CUDA-event interval; the roofline figures need clean ones.
    L.mm_b200_set_serial(1)
    _, _, st_res, prof_res, _, _, _ = timed(True)
    L.mm_b200_set_serial(0)

    out = None
    if rank == 0:
        import json as _json
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = _json.load(open(pk))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        # dominant kernel by device time in the resident timed region
        dom = max(prof_res.items(), key=lambda kv: kv[1][0]) if prof_res else ("none", (0.0, 0))
        dname, (dms, dn) = dom
        step_dev_ms = sum(v[0] for v in prof_res.values())
        # algorithmic HBM bytes of each kernel over the timed region (bytes per unit x units from the live counters)
        nmv, nanch, cells, jobs = st_res.n_minimizers, st_res.n_anchors, st_res.n_dp_cells, st_res.n_dp_jobs
        cells_fast, jobs_fast = st_res.n_dp_cells_fast, st_res.n_dp_jobs_fast
        algo = {
            "k_sketch_count": st_res.n_bases * 0.5, "k_sketch_fill": st_res.n_bases * 0.5 + nmv * 16,
            "k_lookup": nmv * (16 + 16 + 12), "k_fill": nmv * 28 + nanch * (8 + 16), "k_expand": nmv * 28 + nanch * (8 + 16),
            "k_chain_fill": nanch * (16 + 16), "k_chain_tail_warp": nanch * (16 + 16 + 16 + 16),
            "k_ksw": (cells - cells_fast) * 1.0 + (jobs - jobs_fast) * 64, "k_ksw_tpj": cells_fast * 1.0 + jobs_fast * 64,
            "k_encode_reads": st_res.n_bases * 1.5,
            # post-chaining bookkeeping (one fragment per thread): chained anchors in, query + target windows as 4-bit codes, hits out
            "k_post_hits": nanch * 16 + st_res.n_bases * 1.0 + st_res.n_reads * 96, "k_post_align": cells * 0.0 + jobs * 64 + st_res.n_bases * 1.0 + st_res.n_reads * 160,
        }
        a_bytes = algo.get(dname, 0.0)
        achieved = a_bytes / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
        traffic = None
        if os.path.exists(TRAFFIC_JSON):
            traffic = _json.load(open(TRAFFIC_JSON)).get(dname, {}).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "launches": dn, "avg_launch_ms": dms / dn if dn else None,
                "algorithmic_bytes_per_launch": a_bytes / dn if dn else None,
                "share_of_kernel_time": dms / step_dev_ms if step_dev_ms else None,
                "note": "the chaining and DP kernels are integer/latency-bound, not HBM-bound: their HBM fraction is small by construction; "
                        "see k4 (GCUPS against the integer-pipe roofline) and DESIGN.md section 4"}
        ksw_lit_ms = prof_res.get("k_ksw", (0.0, 0))[0]
        ksw_fast_ms = prof_res.get("k_ksw_tpj", (0.0, 0))[0] + prof_res.get("k_ksw_wave", (0.0, 0))[0]
        ksw_ms = ksw_lit_ms + ksw_fast_ms
        lookup_ms = prof_res.get("k_lookup", (0.0, 0))[0]
        gc = lambda c, ms: c / (ms * 1e-3) / 1e9 if ms else None
        k4 = {"gcups": gc(cells, ksw_ms), "gcups_fast_form": gc(cells_fast, ksw_fast_ms), "gcups_literal_form": gc(cells - cells_fast, ksw_lit_ms),
              "cells_per_step": cells / args.steps, "fast_form_cell_share": cells_fast / cells if cells else None,
              "int_roofline_gcups": K4_INT_ROOFLINE_GCUPS, "frac": (gc(cells, ksw_ms) or 0.0) / K4_INT_ROOFLINE_GCUPS,
              "roofline_def": "148 SMs x 128 int lanes x 1.965 GHz / 30 lane-ops per cell (SURVEY.md 8d, one cell per lane-op)"}
        out = {
            "metric": METRIC, "value": reads_res / secs_res, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs_res * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic", "config": workload_config(args.pairs),
            "e2e": {"value": reads_e2e / secs_e2e, "unit": "reads/s", "h2d_bytes_per_step": st_e2e.h2d_bytes // args.steps,
                    "d2h_bytes_per_step": st_e2e.d2h_bytes // args.steps, "ms_per_step": secs_e2e * 1e3 / args.steps},
            "gpu_launches": int(launches_res), "roofline": roof,
            "ksw_gcups": k4["gcups"], "k4": k4,
            "seed_lookup_gbs": algo["k_lookup"] / (lookup_ms * 1e-3) / 1e9 if lookup_ms else None,
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof_res.items(), key=lambda kv: -kv[1][0])},
            "kernel_profile": "CUDA events around every launch in a separate pass of the same K steps with the two shards of a GPU serialised "
                              "(in the timed passes their kernels overlap, which stretches per-kernel intervals)",
            "host_s_per_step": {"seed_chain_call": st_res.t_seedchain / args.steps, "hits": st_res.t_hits / args.steps,
                                "align_host": st_res.t_align_host / args.steps, "dp_call": st_res.t_ksw_total / args.steps,
                                "finish": st_res.t_finish / args.steps, "dp_rounds": st_res.n_dp_rounds / args.steps},
            "host_s_per_step_e2e": {"upload": st_e2e.t_upload / args.steps, "seed_chain_call": st_e2e.t_seedchain / args.steps, "hits": st_e2e.t_hits / args.steps,
                                    "align_host": st_e2e.t_align_host / args.steps, "dp_call": st_e2e.t_ksw_total / args.steps, "finish": st_e2e.t_finish / args.steps,
                                    "total": st_e2e.t_total / args.steps},
            "per_rank_ms_per_step": per_rank_ms, "host_threads": n_threads, "index_build_s": t_index, "index_broadcast_bytes": bcast_bytes, "clocks": clocks_res,
        }
    # ---- CPU baseline next to it (rank 0, N=1 only): the reference fork on a bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        cores = os.cpu_count() or 1
        n_pairs = 200_000
        g1, g2 = make_reads(d, fa, n_pairs, 44, f"sr_ref_{n_pairs}")
        try:
            t = ref_mapping_phase([REF_BIN, "-ax", "sr", "-t", str(cores), fa, g1, g2])
            out["cpu_baseline"] = {"value": 2 * n_pairs / t, "unit": "reads/s", "cores": cores, "kind": "reference",
                                   "sample": f"{n_pairs} pairs (2x150) of the same workload, mapping phase only, -t {cores}"}
        except Exception as e:  # noqa
            out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}
    if cli_result is not None:
        out["cli"] = cli_result
    if rank == 0 and "cpu_baseline" in out:
        pass
    elif rank == 0:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "measured at N=1 only" if world > 1 else "oracle/_ref/minimap2_B not built"}
    L.mm_b200_close_reads(reader)
    L.mm_idx_destroy(mi)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()

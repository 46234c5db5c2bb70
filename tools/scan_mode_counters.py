#!/usr/bin/env python
"""How often does the scan pass over a chunk?  Runs a 1 GB synthetic input through the engine in the three scan modes
(0 sparse, 2 read-first = sparse under -q, 1 dense) and prints mc_scan's counters: chunks, chunks passed over after the
look at their first columns (MC_C_QUIET), lines, kept lines seen by the full parse, records.
usage (GPU box): python tools/scan_mode_counters.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from mcaller_b200 import engine, models, read_qual, synth, synth_device
from mcaller_b200.refindex import ReferenceIndex
reads = 2000
spec = synth.SynthSpec(seed=0, contigs=[("ecoli", 4600000)], n_reads=reads, len_min=1000, len_max=3000)
seqs = {"ecoli": synth.genome(spec, 0).tobytes().decode()}
ref = ReferenceIndex(seqs, "A", motif="GATC", k=6)
meth = {0: (synth.meth_sites(spec, 0, ref.site_fwd_bits[:4600000]), synth.meth_sites(spec, 0, ref.site_rev_bits[:4600000]))}
gen = synth_device.DeviceSynth(spec, ref, meth)
d_text, n, _ = gen.generate(0, reads)
keys, q = synth_device.quality_table_for(spec, 0, reads)
qt = read_qual.build_quality_table(dict(zip(keys, q.tolist())))
model = models.load_model_file(os.path.join(ROOT, "tests", "golden", "models", "r95_twobase_model_NN_6_m6A.pkl"))
dm = models.DeviceModels(model["MH"], model["MG"])
print("first line:", bytes(d_text[:200].cpu().numpy()).split(b"\n")[0])
for qth, dense in ((0.0, None), (13.5, None), (13.5, True)):
    eng = engine.Engine(ref, models=dm, qual_table=qt, skip_thresh=0, qual_thresh=qth, two_models=True, dense=dense)
    res = eng.run_chunk(d_text, n)
    c = res.counters
    print("q=%g mode=%d: chunks=%d quiet=%d lines=%d kept(full parse)=%d records=%d rows=%d" % (qth, eng.scan_mode, (n + 3839) // 3840, c["quiet_chunks"], c["lines"], c["kept"], res.n_records, -1))

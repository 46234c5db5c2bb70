#!/bin/sh
# Builds oracle/_ref/: the UNMODIFIED reference (al-mcintyre/mCaller) laid out so that it can run on a box where
# /root/reference does not exist (the GPU box).  The reference is pure Python, so "building" it means copying its
# modules next to the import shim (Bio.SeqIO / matplotlib / seaborn stand-ins and the sklearn pickle aliases of
# tools/ref_shim).  oracle/_ref/ is git-ignored (reference sources never enter the history) but
# travels with the gpurun snapshot.  Test infrastructure: used by bench.py (--impl reference, cpu_baseline) only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MCALLER_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -f "$REF/mCaller.py" ] || { echo "reference checkout not found at $REF" >&2; exit 1; }
rm -rf "$OUT"
mkdir -p "$OUT/shim"
for f in mCaller.py extract_contexts.py make_bed.py read_qual.py train_model.py load_mCaller_data.py plotlib.py; do
    cp "$REF/$f" "$OUT/$f"
done
cp -r "$HERE/../tools/ref_shim/." "$OUT/shim/"
find "$OUT" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$REF" && sha256sum mCaller.py extract_contexts.py make_bed.py read_qual.py train_model.py load_mCaller_data.py plotlib.py ) > "$OUT/SHA256SUMS"
echo "$OUT"

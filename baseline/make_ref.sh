#!/bin/sh
# Puts the UNMODIFIED Python package of the reference under baseline/_ref/ (git-ignored; it travels to the GPU
# box with the gpurun snapshot) so that bench.py can time the reference's own arcte_worker
# (embedding/arcte/arcte.py:279) on the box's host cores ("cpu_baseline_python").
# `pip install --target baseline/_ref /root/reference` does not work: the reference's setup.py lists packages
# that are not in its tree (eps_randomwalk/benchmarks, setup.py:81-82), so metadata generation fails.  The
# package is pure Python (its Cython build is commented out, setup.py:4-12): a plain copy is what an install
# would have produced.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${1:-/root/reference}"
[ -d "$ref/reveal_graph_embedding" ] || { echo "no reference tree at $ref"; exit 0; }
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
cp -r "$ref/reveal_graph_embedding" "$here/_ref/"
find "$here/_ref" \( -name "*.so" -o -name "*.c" -o -name "*.pyc" \) -type f -exec rm -f {} +   # stale Cython 0.2x output for CPython 3.4
echo "reference package copied to $here/_ref"

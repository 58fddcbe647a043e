#!/usr/bin/env bash
# Stage the UNMODIFIED reference files of the hot path into oracle/_ref/ (git-ignored: never committed; it travels to the GPU
# box with gpurun like the built .so files).  Used only as the checker / timed baseline (oracle/ref_runner.py); nothing
# under gnn_branching_b200/ imports it.  Run in the build container, where /root/reference exists.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/oracle/_ref"
[ -d "$REF/graphnet" ] || { echo "stage_ref: $REF/graphnet not found" >&2; exit 1; }
mkdir -p "$DST/graphnet" "$DST/plnn"
cp "$REF/graphnet/graph_conv.py" "$REF/graphnet/graph_score.py" "$REF/graphnet/graph_score_online.py" "$DST/graphnet/"
cp "$REF/plnn/modules.py" "$DST/plnn/"
( cd "$REF" && sha256sum graphnet/graph_conv.py graphnet/graph_score.py graphnet/graph_score_online.py plnn/modules.py ) > "$DST/SHA256SUMS"
echo "staged $(wc -l < "$DST/SHA256SUMS") reference files into $DST"

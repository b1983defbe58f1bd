#!/bin/bash
# Stage the reference's own hot-path sources, UNMODIFIED, under git-ignored baseline/_ref/ so that the GPU box (which
# has no /root/reference) can time the real CPU path next to the engine: `bench.py --impl reference` and the
# `cpu_baseline` leg import them from there.  baseline/_ref/ is listed in .gitignore (never committed) and not in
# .gpurunignore (it travels with the snapshot like the built .so files).  __graft_entry__.build() runs this when
# /root/reference exists.
set -eu
SRC=${OTHELLO_REFERENCE:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -f "$SRC/self_play_worker.py" ] || { echo "no reference checkout at $SRC"; exit 0; }
mkdir -p "$DST/envs"
for f in MCTS_model.py Models.py self_play_worker.py eval.py envs/__init__.py envs/game.py envs/othello.py; do
  cp -pf "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum MCTS_model.py Models.py self_play_worker.py eval.py envs/__init__.py envs/game.py envs/othello.py ) > "$DST/SHA256SUMS"
echo "staged $(wc -l < "$DST/SHA256SUMS") reference files in $DST"

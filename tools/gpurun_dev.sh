#!/bin/bash
# Development helper: gpurun WITHOUT the 385 MB of shipped checkpoints (the push is charged box time).  The checkpoints are
# moved to a stash outside the repo for the duration of the call and ALWAYS moved back (trap), so the driver's
# round-end snapshot carries checkpoints/_ref/.  KEEP="CRN__ DCCRN__" keeps the files with those prefixes in place.
# usage: [KEEP="prefix ..."] tools/gpurun_dev.sh [--timeout S] -- '<command>'
cd "$(dirname "$0")/.."
STASH=/tmp/ckpt_stash
mkdir -p "$STASH"
restore() { mv "$STASH"/*.pth checkpoints/_ref/ 2>/dev/null; true; }
trap restore EXIT
for f in checkpoints/_ref/*.pth; do
  [ -e "$f" ] || continue
  keep=0
  for k in $KEEP; do case "$(basename "$f")" in "$k"*) keep=1;; esac; done
  [ $keep = 1 ] || mv "$f" "$STASH"/
done
/usr/local/graft/bin/gpurun "$@"

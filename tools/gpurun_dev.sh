#!/bin/bash
# Development helper: gpurun with a temporary .gpurunignore (build objects, optionally more) so that kernel-iteration
# calls push as little as possible -- the push is charged box time.  The ignore file only exists while the call runs,
# so the driver's round-end snapshot is never affected.
# usage: tools/gpurun_dev.sh [--timeout S] -- '<command>'      (extra ignore patterns: GPURUN_DEV_IGNORE="a b c")
cd "$(dirname "$0")/.."
trap 'rm -f .gpurunignore' EXIT
{
  printf 'sixty-years-of-frequency-domain-monaural-speech-enhancement_b200/build/\n'
  for pat in $GPURUN_DEV_IGNORE; do printf '%s\n' "$pat"; done
} > .gpurunignore
/usr/local/graft/bin/gpurun "$@"

#!/bin/sh
# Builds oracle/_ref/bin/HERest_gpu: the reference's HERest with its FBFile call site re-pointed
# at libhfbgpu through bridge/hfbgpu_bridge.c.  HTK's licence forbids redistributing modified
# source, so the four one-line edits are applied to a LOCAL copy under the git-ignored
# oracle/_ref/ (never committed); everything else of HERest -- options, MMF/label/feature I/O,
# -p dump format, M-step -- is the reference's own code, unchanged.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${HTK_REFERENCE:-/root/reference}
OUT=$ROOT/oracle/_ref
CF='-O2 -D_SVID_SOURCE -D_DEFAULT_SOURCE -DOSS_AUDIO -DARCH="x86_64" -w -DPHNALG'
mkdir -p "$OUT/build" "$OUT/bin"
[ -f "$OUT/HTKLib.a" ] || make -C "$ROOT/oracle" ref CC=gcc
SRC=$OUT/build/HERest_gpu.c
sed -e 's|#include "HFB.h"|#include "HFB.h"\n#include "hfbgpu_bridge.h"|' \
    -e 's|if (FBFile(fbInfo, utt, datafn)) {|if (HFBGPU_Queue(fbInfo, utt, datafn)) {|' \
    -e 's|   InitUttInfo(utt, twoDataFiles);|   InitUttInfo(utt, twoDataFiles);\n   if (parMode != 0) HFBGPU_Init(\&hset, fbInfo, pruneInit, pruneInc, pruneLim, minFrwdP, uFlags, trace);|' \
    -e 's|   LoadData(fbInfo->al_hset, utt, dff, datafn, datafn2);|   if (!HFBGPU_FastLoad(utt, datafn, datafn2)) LoadData(fbInfo->al_hset, utt, dff, datafn, datafn2);|' \
    -e 's|   } while (NumArgs()>0);|   } while (NumArgs()>0);\n   if (parMode != 0) HFBGPU_Finish(\&totalT, \&totalPr);|' \
    "$REF/HTKTools/HERest.c" > "$SRC"
grep -q HFBGPU_Queue "$SRC" && grep -q HFBGPU_Init "$SRC" && grep -q HFBGPU_Finish "$SRC" && grep -q HFBGPU_FastLoad "$SRC" || { echo "patch did not apply"; exit 1; }
gcc $CF -I"$REF/HTKLib" -I"$ROOT/include" -I"$ROOT/bridge" -c "$ROOT/bridge/hfbgpu_bridge.c" -o "$OUT/build/hfbgpu_bridge.o"
gcc $CF -ansi -I"$REF/HTKLib" -I"$ROOT/include" -I"$ROOT/bridge" -c "$SRC" -o "$OUT/build/HERest_gpu.o"
gcc -o "$OUT/bin/HERest_gpu" "$OUT/build/HERest_gpu.o" "$OUT/build/hfbgpu_bridge.o" "$OUT/HTKLib.a" \
    -L"$ROOT/htk_b200" -lhfbgpu -lpthread -Wl,-rpath,'$ORIGIN/../../../htk_b200' -lm
echo "built $OUT/bin/HERest_gpu"

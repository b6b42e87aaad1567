#!/bin/bash
# Generates the bitstream corpus with the UNMODIFIED reference encoder (oracle/_ref/TAppEncoderStatic)
# from seeded synthetic YUV (tools/gen_yuv.py) using the cfg files shipped in /root/reference/cfg.
# Usage: tools/make_corpus.sh <name>        (one job; see table below)
#        tools/make_corpus.sh --list
# Outputs corpus/<name>.bin, corpus/<name>.md5 (per-frame MD5 lines printed by TAppDecoderStatic),
# corpus/<name>.yuvmd5 (md5 of the whole decoded YUV).  Raw YUVs live in $TMP_YUV (not in the repo).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
ENC=$ROOT/oracle/_ref/TAppEncoderStatic
DEC=$ROOT/oracle/_ref/TAppDecoderStatic
TMP_YUV=${TMP_YUV:-/tmp/corpus}
mkdir -p "$TMP_YUV" "$ROOT/corpus"
#        name        cfg                                   W    H    frames bd chroma seed extra
JOBS=(
 "c1_intra8_240p     encoder_intra_main.cfg                416  240  16 8  420 1234"
 "s_ra8_240p         encoder_randomaccess_main.cfg         416  240  17 8  420 11"
 "s_ra8_240p_q22     encoder_randomaccess_main.cfg         416  240  17 8  420 12 -q 22"
 "s_ra10_240p        encoder_randomaccess_main10.cfg       416  240  17 10 420 13"
 "s_ld10_240p        encoder_lowdelay_main10.cfg           416  240  9  10 420 14"
 "s_ldp8_240p        encoder_lowdelay_P_main.cfg           416  240  9  8  420 15"
 "s_intra10_240p_q22 encoder_intra_main10.cfg              416  240  4  10 420 16 -q 22"
 "s_rext444_240p     encoder_intra_high_throughput_rext.cfg 416 240  3  12 444 17 --InternalBitDepth=12"
 "s_ra8_odd          encoder_randomaccess_main.cfg         200  136  9  8  420 18 -q 27"
 "s_ra422_240p       encoder_randomaccess_main_rext.cfg    416  240  9  10 422 19 --InternalBitDepth=10 -q 27"
 "s_sl8_240p         encoder_randomaccess_main.cfg         416  240  9  8  420 20 --ScalingList=1 -q 27"
 "s_pcm_240p         encoder_randomaccess_main10.cfg       416  240  5  8  420 21 --InternalBitDepth=10 --PCMEnabledFlag=1 --PCMFilterDisableFlag=1 -q 1"
 "s_lossless_240p    encoder_randomaccess_main.cfg         208  120  5  8  420 22 --TransquantBypassEnableFlag=1 --CUTransquantBypassFlagForce=1 -q 30"
 "s_tiles_240p       encoder_randomaccess_main.cfg         640  256  9  8  420 27 --TileUniformSpacing=1 --NumTileColumnsMinus1=1 --NumTileRowsMinus1=1 --LFCrossTileBoundaryFlag=0 -q 27"
 "s_wavefront_240p   encoder_randomaccess_main.cfg         416  240  9  8  420 28 --WaveFrontSynchro=1 -q 27"
 "s_cip_240p         encoder_randomaccess_main.cfg         416  240  9  8  420 29 --ConstrainedIntraPred=1 -q 30"
 "s_ctu32_240p       encoder_randomaccess_main.cfg         416  240  9  8  420 31 --MaxCUWidth=32 --MaxCUHeight=32 --MaxPartitionDepth=3 -q 27"
 "s_ctu16_240p       encoder_lowdelay_P_main.cfg           416  240  9  8  420 32 --MaxCUWidth=16 --MaxCUHeight=16 --MaxPartitionDepth=2 --QuadtreeTULog2MaxSize=4 -q 27"
 "s_dqp_240p         encoder_randomaccess_main.cfg         416  240  9  8  420 33 --MaxDeltaQP=3 --MaxCuDQPDepth=2 --CbQpOffset=3 --CrQpOffset=-2 --LoopFilterOffsetInPPS=1 --LoopFilterBetaOffset_div2=2 --LoopFilterTcOffset_div2=-1 -q 30"
 "s_ra444_240p       encoder_randomaccess_main_rext.cfg    416  240  9  8  444 34 --InternalBitDepth=8 -q 27"
 "s_nolf_240p        encoder_randomaccess_main.cfg         416  240  9  8  420 35 --SAO=0 --LoopFilterDisable=1 -q 30"
 "s_tiles2_240p      encoder_randomaccess_main.cfg         640  256  9  8  420 36 --TileUniformSpacing=1 --NumTileColumnsMinus1=1 --NumTileRowsMinus1=1 --LFCrossTileBoundaryFlag=1 -q 30"
 "s_wpp_240p         encoder_lowdelay_P_main.cfg           416  240  9  8  420 23 --WeightedPredP=1 -q 30"
 "s_wpb_240p         encoder_randomaccess_main.cfg         416  240  9  8  420 24 --WeightedPredB=1 --WeightedPredP=1 -q 30"
 "c2_ra8_1080p       encoder_randomaccess_main.cfg         1920 1080 64 8  420 2"
 "c3_ra10_2160p      encoder_randomaccess_main10.cfg       3840 2160 33 10 420 3"
 "c4_rext444_1080p   encoder_intra_high_throughput_rext.cfg 1920 1080 8 12 444 4 --InternalBitDepth=12"
 "c5_ld10_2160p_s50  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 50"
 "c5_ld10_2160p_s51  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 51"
 "c5_ld10_2160p_s52  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 52"
 "c5_ld10_2160p_s53  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 53"
 "c5_ld10_2160p_s54  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 54"
 "c5_ld10_2160p_s55  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 55"
 "c5_ld10_2160p_s56  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 56"
 "c5_ld10_2160p_s57  encoder_lowdelay_main10.cfg           3840 2160 17 10 420 57"
 "m_ra10_1080p       encoder_randomaccess_main10.cfg       1920 1080 17 10 420 5"
 "f_ra10_1080p       encoder_randomaccess_main10.cfg       1920 1080 17 10 420 6"
 "f_ra10_2160p       encoder_randomaccess_main10.cfg       3840 2160 33 10 420 7"
 "q27_ra10_2160p     encoder_randomaccess_main10.cfg       3840 2160 33 10 420 3 -q 27"
 "q22_ra8_1080p      encoder_randomaccess_main.cfg         1920 1080 33 8  420 2 -q 22"
 "s_crc_240p         encoder_randomaccess_main.cfg         416  240  9  8  420 41 --SEIDecodedPictureHash=2 -q 30"
 "s_slseg_240p       encoder_randomaccess_main.cfg         416  240  9  8  420 43 --SliceSegmentMode=1 --SliceSegmentArgument=7 -q 30"
 "s_gray400_240p     encoder_randomaccess_main_rext.cfg    416  240  5  8  400 44 --InternalBitDepth=8 -q 30"
 "s_cksum_240p       encoder_randomaccess_main10.cfg       416  240  9  10 420 42 --SEIDecodedPictureHash=3 -q 30"
 "s_mintu8_240p      encoder_intra_main.cfg                416  240  2  8  420 45 --QuadtreeTULog2MinSize=3 --MaxPartitionDepth=3 -q 32"
 "s_cra_240p         encoder_randomaccess_main.cfg         416  240  25 8  420 46 --IntraPeriod=8 -q 30"
 "s_ramintu8_240p    encoder_randomaccess_main10.cfg       416  240  9  10 420 47 --QuadtreeTULog2MinSize=3 --MaxPartitionDepth=3 -q 27"
 "s_mintu16_240p     encoder_randomaccess_main.cfg         416  256  5  8  420 48 --QuadtreeTULog2MinSize=4 --MaxPartitionDepth=2 -q 27"
 "s_mintu32_240p     encoder_randomaccess_main.cfg         448  256  5  8  420 49 --QuadtreeTULog2MinSize=5 --MaxPartitionDepth=1 --QuadtreeTUMaxDepthInter=1 --QuadtreeTUMaxDepthIntra=1 -q 27"
)
WANT=$1
# s_switch_240p: five coded video sequences back to back, every one starting with an IDR and four of the five activating an SPS of
# another geometry (416x240 -> 200x136 -> 416x240 -> 208x120 lossless -> 416x240): resolution switches with DPB flushes in between.
if [ "$WANT" == "s_switch_240p" ]; then
  out=$ROOT/corpus/s_switch_240p
  for s in s_ra8_240p s_ra8_odd s_ra8_240p_q22 s_lossless_240p; do [ -s "$ROOT/corpus/$s.bin" ] || "$0" $s; done
  cat "$ROOT/corpus/s_ra8_240p.bin" "$ROOT/corpus/s_ra8_odd.bin" "$ROOT/corpus/s_ra8_240p_q22.bin" "$ROOT/corpus/s_lossless_240p.bin" "$ROOT/corpus/s_ra8_240p.bin" > "$out.bin"
  "$DEC" -b "$out.bin" -d 0 -o "$TMP_YUV/s_switch.dec.yuv" > "$out.dec.log" 2>&1
  grep -o 'POC.*' "$out.dec.log" | sed -E 's/\[DT +[0-9.]+\] //' > "$out.md5"
  md5sum < "$TMP_YUV/s_switch.dec.yuv" | awk '{print $1}' > "$out.yuvmd5"
  echo "s_switch_240p: done ($(stat -c %s "$out.bin") bytes, $(grep -c OK "$out.md5") pictures OK)"
  exit 0
fi
# s_seek_240p: a caller that seeks — the parameter sets of s_cra_240p followed by everything from its SECOND CRA picture on (open GOP:
# the CRA's leading RASL pictures reference pictures that were never sent; the decoder must drop them, TDecTop::isRandomAccessSkipPicture,
# TDecTop.cpp:1112-1160).  The pin is what the unmodified TAppDecoder prints for this cut.
if [ "$WANT" == "s_seek_240p" ]; then
  out=$ROOT/corpus/s_seek_240p
  [ -s "$ROOT/corpus/s_cra_240p.bin" ] || "$0" s_cra_240p
  python3 - "$ROOT/corpus/s_cra_240p.bin" "$out.bin" <<'PY'
import re, sys
d = open(sys.argv[1], 'rb').read()
pos = [m.start() for m in re.finditer(b'\x00\x00\x01', d)]
pos = [p - 1 if p > 0 and d[p - 1] == 0 else p for p in pos]                 # keep 4-byte start codes whole
nals = [(p, pos[i + 1] if i + 1 < len(pos) else len(d)) for i, p in enumerate(pos)]
typ = lambda n: (d[n[0] + (4 if d[n[0] + 2] == 0 else 3)] >> 1) & 0x3f
cras = [i for i, n in enumerate(nals) if typ(n) == 21]
assert len(cras) >= 2, "need two CRA pictures"
start = cras[1]
while start > 0 and typ(nals[start - 1]) in (35, 39): start -= 1            # AUD / prefix SEI of the CRA's access unit (the suffix SEI before it belongs to the previous picture)
out = b''.join(d[a:b] for i, (a, b) in enumerate(nals) if typ((a, b)) in (32, 33, 34) and i < start)
out += d[nals[start][0]:]
open(sys.argv[2], 'wb').write(out)
PY
  "$DEC" -b "$out.bin" -d 0 -o "$TMP_YUV/s_seek.dec.yuv" > "$out.dec.log" 2>&1
  grep -o 'POC.*' "$out.dec.log" | sed -E 's/\[DT +[0-9.]+\] //' > "$out.md5"
  md5sum < "$TMP_YUV/s_seek.dec.yuv" | awk '{print $1}' > "$out.yuvmd5"
  echo "s_seek_240p: done ($(stat -c %s "$out.bin") bytes, $(grep -c OK "$out.md5") pictures OK)"
  exit 0
fi
# s_lost_240p: s_ra8_240p without its second coded picture (POC 8, a reference of everything that follows): the decoder conceals it
# (TDecTop::xCreateLostPicture, TDecTop.cpp:233-281: a copy of the closest picture stands in and is output as well).  The pin is what the
# unmodified TAppDecoder prints for this cut: 16 status lines, the SEI MD5s of the pictures that depend on the lost one mismatch by design.
if [ "$WANT" == "s_lost_240p" ]; then
  out=$ROOT/corpus/s_lost_240p
  [ -s "$ROOT/corpus/s_ra8_240p.bin" ] || "$0" s_ra8_240p
  python3 - "$ROOT/corpus/s_ra8_240p.bin" "$out.bin" <<'PY'
import re, sys
d = open(sys.argv[1], 'rb').read()
pos = [m.start() for m in re.finditer(b'\x00\x00\x01', d)]
pos = [p - 1 if p > 0 and d[p - 1] == 0 else p for p in pos]
nals = [(p, pos[i + 1] if i + 1 < len(pos) else len(d)) for i, p in enumerate(pos)]
typ = lambda n: (d[n[0] + (4 if d[n[0] + 2] == 0 else 3)] >> 1) & 0x3f
vcl = [i for i, n in enumerate(nals) if typ(n) < 32]
drop = {vcl[1]}
if typ(nals[vcl[1] + 1]) == 40: drop.add(vcl[1] + 1)                      # its decoded-picture-hash SEI
open(sys.argv[2], 'wb').write(b''.join(d[a:b] for i, (a, b) in enumerate(nals) if i not in drop))
PY
  "$DEC" -b "$out.bin" -d 0 -o "$TMP_YUV/s_lost.dec.yuv" > "$out.dec.log" 2>&1 || true
  grep -o 'POC.*' "$out.dec.log" | sed -E 's/\[DT +[0-9.]+\] //' > "$out.md5"
  md5sum < "$TMP_YUV/s_lost.dec.yuv" | awk '{print $1}' > "$out.yuvmd5"
  echo "s_lost_240p: done ($(stat -c %s "$out.bin") bytes, $(grep -c POC "$out.md5") status lines, $(grep -c OK "$out.md5") OK)"
  exit 0
fi
if [ "$1" == "--list" ]; then for j in "${JOBS[@]}"; do echo "$j" | awk '{print $1}'; done; exit 0; fi
for j in "${JOBS[@]}"; do
  set -- $j
  name=$1; cfg=$2; W=$3; H=$4; F=$5; BD=$6; CH=$7; SEED=$8; shift 8; EXTRA="$*"
  [ "$name" == "$WANT" ] || continue
  out=$ROOT/corpus/$name
  [ -s "$out.bin" ] && [ -s "$out.yuvmd5" ] && { echo "$name: exists"; exit 0; }
  yuv=$TMP_YUV/$name.yuv
  if [ "$name" == "s_wpp_240p" ] || [ "$name" == "s_wpb_240p" ]; then python "$ROOT/tools/gen_fade_yuv.py" "$yuv" --width $W --height $H --frames $F --seed $SEED   # fade: non-trivial WP weights
  elif [ "$name" == "s_pcm_240p" ]; then python "$ROOT/tools/gen_pcm_yuv.py" "$yuv" --width $W --height $H --frames $F --seed $SEED   # noise: makes the encoder choose I_PCM
  elif [ "${name:0:2}" == "f_" ]; then python "$ROOT/tools/gen_yuv.py" "$yuv" --width $W --height $H --frames $F --bitdepth $BD --seed $SEED --chroma $CH --pan 2.75,1.25   # quarter-sample pan: fractional motion vectors
  else python "$ROOT/tools/gen_yuv.py" "$yuv" --width $W --height $H --frames $F --bitdepth $BD --seed $SEED --chroma $CH; fi
  CF=""; [ "$CH" != "420" ] && CF="--InputChromaFormat=$CH"
  "$ENC" -c "$REF/cfg/$cfg" -i "$yuv" -wdt $W -hgt $H -f $F -fr 30 --InputBitDepth=$BD $CF \
      --SEIDecodedPictureHash=1 $EXTRA -b "$out.bin.tmp" -o "$TMP_YUV/$name.rec.yuv" > "$out.enc.log" 2>&1
  mv "$out.bin.tmp" "$out.bin"
  "$DEC" -b "$out.bin" -d 0 -o "$TMP_YUV/$name.dec.yuv" > "$out.dec.log" 2>&1
  cmp "$TMP_YUV/$name.rec.yuv" "$TMP_YUV/$name.dec.yuv"
  grep -o 'POC.*' "$out.dec.log" | sed -E 's/\[DT +[0-9.]+\] //' > "$out.md5"
  md5sum < "$TMP_YUV/$name.dec.yuv" | awk '{print $1}' > "$out.yuvmd5"
  rm -f "$yuv" "$TMP_YUV/$name.rec.yuv"
  echo "$name: done ($(stat -c %s "$out.bin") bytes)"
  exit 0
done
echo "unknown job $WANT"; exit 1

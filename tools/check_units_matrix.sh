#!/bin/bash
# Partition units of 8 samples (smallest transform block 8) combined with the other tools, on the CPU (build container only): the
# reference encoder makes a 5-picture stream per combination, frontend/_build/hmdec_cli_verify runs HM's own CPU reconstruction on our
# parse (every SEI MD5 must be OK) and dumps the records, tools/check_oracle.py requires the oracle to reproduce HM's MD5s after CU
# reconstruction, deblocking and SAO from those records, and the records must not change with HM's own BS / availability routines.
# usage: tools/check_units_matrix.sh        (about a minute; nothing is written into the repo)
ROOT=$(cd "$(dirname "$0")/.." && pwd); REF=${REF:-/root/reference}; T=${TMPDIR:-/tmp}/units_matrix; mkdir -p $T; rc=0
run() {  # name cfg W H bitdepth chroma extra...
  local name=$1 cfg=$2 W=$3 H=$4 BD=$5 CH=$6; shift 6
  python "$ROOT/tools/gen_yuv.py" $T/$name.yuv --width $W --height $H --frames 5 --bitdepth $BD --seed 64 --chroma $CH
  local CF=""; [ "$CH" != "420" ] && CF="--InputChromaFormat=$CH --InternalBitDepth=$BD"
  "$ROOT/oracle/_ref/TAppEncoderStatic" -c $REF/cfg/$cfg -i $T/$name.yuv -wdt $W -hgt $H -f 5 -fr 30 --InputBitDepth=$BD $CF --SEIDecodedPictureHash=1 \
      --QuadtreeTULog2MinSize=3 --MaxPartitionDepth=3 "$@" -q 30 -b $T/$name.bin -o $T/$name.rec.yuv > $T/$name.enc.log 2>&1 || { echo "$name: encoder failed"; rc=1; return; }
  local ok=$("$ROOT/frontend/_build/hmdec_cli_verify" -b $T/$name.bin --dump $T/$name.hmr 2>&1 | grep -c "(OK)")
  local orc=$(cd "$ROOT" && python tools/check_oracle.py $T/$name.hmr | tail -1 | awk '{print $NF}')
  HMDEC_B200_BS_FLAGS=1 HMDEC_B200_HM_AVAIL=1 "$ROOT/frontend/_build/hmdec_cli_verify" -b $T/$name.bin --dump $T/$name.ab.hmr > /dev/null 2>&1
  local same=differs; cmp -s $T/$name.hmr $T/$name.ab.hmr && same=identical
  echo "$name: HM reconstruction on our parse $ok/5 OK, oracle vs HM stage MD5s $orc, records with HM's BS / availability routines $same"
  [ "$ok" == "5" ] && [ "$orc" == "OK" ] && [ "$same" == "identical" ] || rc=1
  rm -f $T/$name.yuv $T/$name.rec.yuv
}
run u8_422      encoder_randomaccess_main_rext.cfg 416 240 10 422
run u8_444      encoder_randomaccess_main_rext.cfg 416 240 8  444
run u8_tiles    encoder_randomaccess_main.cfg      640 256 8  420 --TileUniformSpacing=1 --NumTileColumnsMinus1=1 --NumTileRowsMinus1=1 --LFCrossTileBoundaryFlag=0
run u8_cip      encoder_randomaccess_main.cfg      640 256 8  420 --ConstrainedIntraPred=1
run u8_wavefront encoder_randomaccess_main.cfg     640 256 8  420 --WaveFrontSynchro=1
run u8_slseg    encoder_randomaccess_main.cfg      640 256 8  420 --SliceSegmentMode=1 --SliceSegmentArgument=7
run u8_ldp      encoder_lowdelay_P_main.cfg        416 240 8  420
run u8_sl       encoder_randomaccess_main.cfg      416 240 8  420 --ScalingList=1
exit $rc

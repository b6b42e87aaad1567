#!/bin/bash
# Regenerates tests/golden/ from corpus/ (needs oracle/_ref and frontend/_build; container only).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
for s in c1_intra8_240p s_ra8_240p s_ra8_240p_q22 s_ra10_240p s_ld10_240p s_ldp8_240p s_intra10_240p_q22 s_rext444_240p s_ra8_odd s_ra422_240p s_sl8_240p s_pcm_240p s_lossless_240p s_wpp_240p s_wpb_240p s_tiles_240p s_wavefront_240p s_cip_240p s_ctu32_240p s_ctu16_240p s_dqp_240p s_ra444_240p s_nolf_240p s_tiles2_240p s_crc_240p s_cksum_240p s_slseg_240p s_switch_240p s_gray400_240p s_cra_240p s_seek_240p s_mintu8_240p s_ramintu8_240p s_mintu16_240p s_mintu32_240p s_lost_240p; do
  [ -s "$ROOT/corpus/$s.bin" ] || "$ROOT/tools/make_corpus.sh" $s
  "$ROOT/frontend/_build/hmdec_cli_verify" -b "$ROOT/corpus/$s.bin" --dump /tmp/g_$s.hmr > /dev/null || [ $s == s_lost_240p ]     # (s_lost: hash mismatches by design, exit status 1)
  gzip -9 -c /tmp/g_$s.hmr > "$ROOT/tests/golden/$s.hmr.gz"
  cp "$ROOT/corpus/$s.bin" "$ROOT/corpus/$s.md5" "$ROOT/tests/golden/"
done

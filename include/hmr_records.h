/* hmr_records.h — flat per-frame record formats that cross the host->device boundary.
 *
 * The host side (HM's serial CABAC parser, kept in C++ by design) resolves ALL of HM's indexing
 * (z-order partitions, TComTU recursion, neighbour availability, QP derivation, boundary strength,
 * SAO merge) and emits these plain-old-data arrays once per picture.  The reconstruction engine
 * (libhmrecon.so, CUDA sm_100a) and the CPU oracle (oracle/hm_oracle.c) consume exactly the same
 * arrays.  Every field cites the reference code that defines its meaning (paths relative to
 * /root/reference/source/Lib).
 *
 * All coordinates are in samples of the component the record refers to, relative to the picture
 * origin.  All structs are little-endian, naturally aligned, no implicit padding.
 */
#ifndef HMR_RECORDS_H
#define HMR_RECORDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMR_MAGIC        0x52524d48u /* "HMRR" */
#define HMR_VERSION      4u
#define HMR_MAX_SLOTS    16          /* DPB slots addressable by a PU record (HM grows the DPB on demand, TDecTop.cpp:180-186) */
#define HMR_NO_OFFSET    0xffffffffu

/* chroma_format values = HM ChromaFormat enum (TypeDef.h CHROMA_400..CHROMA_444) */
enum { HMR_CHROMA_400 = 0, HMR_CHROMA_420 = 1, HMR_CHROMA_422 = 2, HMR_CHROMA_444 = 3 };

/* hmr_frame_hdr.flags */
enum {
  HMR_FRM_STRONG_INTRA_SMOOTHING = 1u << 0, /* SPS strong_intra_smoothing (TComPattern.cpp:197-214) */
  HMR_FRM_DEBLOCK               = 1u << 1, /* at least one slice has deblocking enabled; BS map is valid */
  HMR_FRM_SAO                   = 1u << 2, /* SPS SAO on and at least one CTU/component not OFF (TComSampleAdaptiveOffset.cpp:717-724) */
  HMR_FRM_HAS_NOFILTER          = 1u << 3, /* pcm_loop_filter_disabled or cu_transquant_bypass present: cu_flags map is valid */
  HMR_FRM_HAS_CCP               = 1u << 4, /* PPS cross_component_prediction (4:4:4): luma residuals are kept for chroma TUs */
  HMR_FRM_IS_REFERENCE          = 1u << 5, /* informational */
  HMR_FRM_INTRA_ONLY            = 1u << 6, /* informational: no PU records */
  HMR_FRM_SCALING_LIST          = 1u << 7, /* scaling lists in use: hmr_frame_desc.scaling is valid (TComTrQuant.cpp:1230-1276) */
  HMR_FRM_WEIGHTED_PRED         = 1u << 8, /* explicit weighted prediction applies to this picture's slices (TComSlice::applyWP):
                                              hmr_frame_desc.wp and .pu_refidx are valid */
};

/* Scaling factors m[y][x] (1..255) of the active scaling list, expanded to the transform size the way
 * TComTrQuant::processScalingListDec does (TComTrQuant.cpp:3092-3106: 16x16/32x32 replicate the 8x8 list, DC replaced):
 * for size s = log2(N)-2 and list l = 3*(inter) + component: N*N bytes at HMR_SCALING_OFFSET(s) + l*N*N, raster like the levels.
 * The dequantiser multiplies by g_invQuantScales[qp%6] * m instead of g_invQuantScales[qp%6] << 4. */
#define HMR_SCALING_OFFSET(s) ((s) == 0 ? 0 : (s) == 1 ? 96 : (s) == 2 ? 480 : 2016)
#define HMR_SCALING_BYTES     8160

typedef struct hmr_frame_hdr {
  uint32_t magic;            /* HMR_MAGIC */
  uint32_t version;          /* HMR_VERSION */
  int32_t  width, height;    /* luma samples, full coded size (SPS pic_width/height_in_luma_samples) */
  int32_t  poc;
  uint8_t  chroma_format;    /* HMR_CHROMA_* */
  uint8_t  bit_depth_luma;   /* g_bitDepth[CHANNEL_TYPE_LUMA]   (TComRom.cpp:319) */
  uint8_t  bit_depth_chroma; /* g_bitDepth[CHANNEL_TYPE_CHROMA] */
  uint8_t  log2_ctu;         /* log2(g_uiMaxCUWidth), 4..6 */
  uint8_t  out_slot;         /* DPB slot this picture is reconstructed into */
  uint8_t  slice_type;       /* 0=B 1=P 2=I of the first slice (informational) */
  int8_t   pps_cb_qp_offset; /* deblock chroma QP uses the PPS offset only (TComLoopFilter.cpp:759) */
  int8_t   pps_cr_qp_offset;
  uint32_t flags;            /* HMR_FRM_* */
  uint32_t n_tu;             /* residual records  (hmr_tu)    */
  uint32_t n_coef;           /* int16 coefficient entries (multiple of 16) */
  uint32_t n_intra;          /* intra records     (hmr_intra) */
  uint32_t n_pu;             /* inter records     (hmr_pu)    */
  uint32_t n_mc_tiles;       /* total 16x16-luma tiles over all PUs == pu_tile_prefix[n_pu] */
  uint32_t n_ctu;            /* CTUs in the picture, raster order */
  uint32_t tu_first[5];      /* hmr_tu records are grouped by size: [tu_first[k], tu_first[k+1]) hold log2_size == k+2 */
  uint32_t reserved[1];
} hmr_frame_hdr;             /* 80 bytes */

/* ---- residual (dequant + inverse transform) records: TComTrQuant::invTransformNxN (TComTrQuant.cpp:1423-1548) ---- */
enum {
  HMR_TU_CODED    = 1u << 0, /* cbf != 0: coefficients present at coef_off (else only CCP contributes) */
  HMR_TU_INTRA    = 1u << 1, /* residual is consumed by the intra stage (kept in the compact residual buffer) */
  HMR_TU_DST      = 1u << 2, /* 4x4 luma intra: DST-VII (TComTU.cpp:218-224, TComTrQuant.cpp:437-461) */
  HMR_TU_TSKIP    = 1u << 3, /* transform_skip_flag (TComTrQuant.cpp:1920-1959) */
  HMR_TU_BYPASS   = 1u << 4, /* cu_transquant_bypass (TComTrQuant.cpp:1475-1487) */
  HMR_TU_ROTATE   = 1u << 5, /* isNonTransformedResidualRotated (TComTU.cpp:227-233) */
  HMR_TU_RDPCM_H  = 1u << 6, /* invRdpcmNxN horizontal (TComTrQuant.cpp:1737-1792) */
  HMR_TU_RDPCM_V  = 1u << 7, /* invRdpcmNxN vertical   */
};

typedef struct hmr_tu {
  uint16_t x, y;        /* top-left, component samples */
  uint8_t  comp;        /* 0=Y 1=Cb 2=Cr */
  uint8_t  log2_size;   /* 2..5 (square; 4:2:2 N x 2N chroma TUs are emitted as two squares, TComTrQuant.cpp:1437-1464) */
  uint8_t  flags;       /* HMR_TU_* */
  uint8_t  qp;          /* QpParam::Qp = per*6+rem (TComTrQuant.cpp:71-119) */
  int8_t   ccp_alpha;   /* cross-component prediction alpha, 0 = off (TComTrQuant.cpp:3294-3335) */
  uint8_t  pad[3];
  uint32_t coef_off;    /* offset (int16 units) of this TU's N*N levels (raster, stride N) in the coefficient buffer;
                           the residual of INTRA TUs (and of luma TUs when HMR_FRM_HAS_CCP) is written to the same
                           offset of the compact residual buffer */
  uint32_t luma_off;    /* CCP only: compact-residual offset of the co-located luma TU, or HMR_NO_OFFSET */
} hmr_tu;               /* 20 bytes */

/* ---- intra prediction records, in HM decode order, grouped per CTU and per component:
 *      TDecCu::xIntraRecBlk (TDecCu.cpp:483-659), TComPattern.cpp:107-520, TComPrediction.cpp:182-491,746-835 ---- */
enum {
  HMR_INTRA_FILTER_REFS  = 1u << 0, /* filteringIntraReferenceSamples() result (TComPattern.cpp:531-556) */
  HMR_INTRA_AVAIL_CORNER = 1u << 1, /* above-left unit available (TComPattern.cpp:140) */
  HMR_INTRA_NO_EDGE_FLT  = 1u << 2, /* implicit RDPCM + bypass disables DC/H/V boundary filters (TComPrediction.cpp:476) */
  HMR_INTRA_LUMA_RULES   = 1u << 3, /* channel is luma: DC/edge filters and strong smoothing may apply */
};

/* I_PCM block (TDecCu::xReconPCM / xDecodePCMTexture, TDecCu.cpp:771-842): "prediction" 0, and the block's residual record
 * (HMR_TU_BYPASS) carries the PCM samples already shifted to the internal bit depth.  Keeps PCM inside the decode order. */
#define HMR_INTRA_MODE_PCM 35

typedef struct hmr_intra {
  uint16_t x, y;        /* top-left, component samples */
  uint8_t  comp;
  uint8_t  log2_size;   /* 2..5 */
  uint8_t  mode;        /* final prediction mode 0..34 after DM / 4:2:2 mapping (TDecCu.cpp:524-526), or HMR_INTRA_MODE_PCM */
  uint8_t  flags;       /* HMR_INTRA_* */
  /* Neighbour availability per "unit" (= 4 luma samples = 4>>csx chroma samples, TComPattern.cpp:119-127),
     bit i = unit i counted AWAY from the top-left corner: left/below-left downwards, above/above-right rightwards. */
  uint8_t  avail_left, avail_below_left, avail_above, avail_above_right;
  uint32_t resid_off;   /* compact-residual offset (== hmr_tu.coef_off of the same block) or HMR_NO_OFFSET when no residual */
} hmr_intra;            /* 16 bytes */

/* per CTU, per component: [first, first+count) into the hmr_intra array */
typedef struct hmr_ctu_intra_range {
  uint32_t first[3];
  uint32_t count[3];
} hmr_ctu_intra_range;  /* 24 bytes */

/* ---- inter prediction records: TComPrediction::motionCompensation (TComPrediction.cpp:514-698) ---- */
enum { HMR_PU_L0 = 1u, HMR_PU_L1 = 2u };

typedef struct hmr_pu {
  uint16_t x, y;        /* luma samples */
  uint8_t  w, h;        /* luma samples, 4..64 */
  uint8_t  lists;       /* HMR_PU_L0|HMR_PU_L1 after xCheckIdenticalMotion (TComPrediction.cpp:497-512) */
  uint8_t  slots;       /* DPB slot of the list-0 reference in bits 0-3, of the list-1 reference in bits 4-7 */
  int16_t  mv[2][2];    /* [list][x,y] quarter-luma-sample, ALREADY clipped by TComDataCU::clipMv (TComDataCU.cpp:3102-3114) */
} hmr_pu;               /* 16 bytes */

/* ---- explicit weighted prediction: TComWeightPrediction::getWpScaling / addWeightUni / addWeightBi
 *      (TComWeightPrediction.cpp:44-53,75-196,211-286).  One entry per (list, refIdx, component); a PU finds its entries
 *      through pu_refidx[pu] = refIdx of list 0 in bits 0-3, of list 1 in bits 4-7. */
typedef struct hmr_wp {
  int16_t  weight;      /* iWeight */
  int16_t  offset;      /* iOffset already multiplied by the offset scaling factor (1 << (bitDepth-8) unless high_precision_offsets) */
  uint8_t  log2_denom;  /* uiLog2WeightDenom */
  uint8_t  pad;
} hmr_wp;               /* 6 bytes; table = hmr_wp[2][16][3] */
#define HMR_WP_ENTRIES (2 * 16 * 3)

/* ---- per-CTU side info: SAO (TComSampleAdaptiveOffset.cpp:375-714) and slice-level deblock offsets ---- */
enum { HMR_SAO_OFF = 0, HMR_SAO_EO_0 = 1, HMR_SAO_EO_90 = 2, HMR_SAO_EO_135 = 3, HMR_SAO_EO_45 = 4, HMR_SAO_BO = 5 };
/* hmr_ctu.avail bits: TComPicSym::deriveLoopFilterBoundaryAvailibility (TComPicSym.cpp:365-460) */
enum { HMR_AV_L = 1, HMR_AV_R = 2, HMR_AV_A = 4, HMR_AV_B = 8, HMR_AV_AL = 16, HMR_AV_AR = 32, HMR_AV_BL = 64, HMR_AV_BR = 128 };

typedef struct hmr_sao {
  uint8_t  type;        /* HMR_SAO_* */
  uint8_t  band;        /* BO: first band (typeAuxInfo) */
  int16_t  off[4];      /* EO: offsets for edgeIdx -2,-1,+1,+2 ; BO: offsets of bands band..band+3 (mod 32); already scaled */
} hmr_sao;              /* 10 bytes */

typedef struct hmr_ctu {
  hmr_sao  sao[3];
  uint8_t  avail;       /* HMR_AV_* */
  int8_t   beta_offset_div2; /* slice deblocking offsets of the slice containing this CTU (TComLoopFilter.cpp:565-566) */
  int8_t   tc_offset_div2;
  uint8_t  pad[3];
} hmr_ctu;              /* 36 bytes */

/* ---- dense maps ----
 * bs map   : uint8 per 4x4 luma unit, raster, stride = ceil(width/4):
 *            bits 0-1 = BS of the VERTICAL edge on the unit's left boundary,
 *            bits 2-3 = BS of the HORIZONTAL edge on the unit's top boundary
 *            (only units on the 8x8 luma grid carry non-zero values, TComLoopFilter.cpp:199-215)
 * qp map   : int8 per 8x8 luma block (min CU), raster, stride = ceil(width/8): TComDataCU::getQP
 * cu_flags : uint8 per 8x8 luma block: bit0 = "no loop filter" (IPCM with pcm_loop_filter_disabled, or lossless)
 */
enum { HMR_CU_NOFILTER = 1 };

/* One picture's worth of host-side arrays (what hmr_submit_frame takes). */
typedef struct hmr_frame_desc {
  const hmr_frame_hdr*       hdr;
  const hmr_tu*              tu;              /* [n_tu] */
  const int16_t*             coef;            /* [n_coef] */
  const hmr_intra*           intra;           /* [n_intra] */
  const hmr_ctu_intra_range* intra_range;     /* [n_ctu] */
  const hmr_pu*              pu;              /* [n_pu] */
  const uint32_t*            pu_tile_prefix;  /* [n_pu+1] exclusive prefix sum of ceil(w/16)*ceil(h/16) */
  const hmr_ctu*             ctu;             /* [n_ctu] */
  const uint8_t*             bs;              /* [(W/4)*(H/4)] or NULL when !HMR_FRM_DEBLOCK */
  const int8_t*              qp;              /* [(W/8)*(H/8)] */
  const uint8_t*             cu_flags;        /* [(W/8)*(H/8)] or NULL when !HMR_FRM_HAS_NOFILTER */
  const uint8_t*             scaling;         /* [HMR_SCALING_BYTES] or NULL when !HMR_FRM_SCALING_LIST */
  const hmr_wp*              wp;              /* [HMR_WP_ENTRIES] or NULL when !HMR_FRM_WEIGHTED_PRED */
  const uint8_t*             pu_refidx;       /* [n_pu] or NULL when !HMR_FRM_WEIGHTED_PRED */
} hmr_frame_desc;

#ifdef __cplusplus
}
#endif
#endif /* HMR_RECORDS_H */

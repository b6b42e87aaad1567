// common.cuh — device-side view of one picture's job (shared by all kernels of libhmrecon.so)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hmr_records.h"

#define HMR_INTRA_JOB_WORDS 16      // 512 (row, component) jobs: 170 CTU rows

// Order in which intra_kernel's CTAs claim the (CTU row, component) jobs of a picture = bit order of intra_job_mask.  Per component the
// rows ascend (a job only ever waits for the row above of its own component, which must have been handed out before it).  Between the
// components luma goes first: a luma row takes about twice as long as a chroma row and its wavefront is the critical path of the picture,
// so a chroma pair follows every SECOND luma row (Y0 Cb0 Cr0 Y1 | Y2 Cb1 Cr1 Y3 | ...) and the second half of the chroma rows comes
// after the last luma row — with the launcher's CTA limit, row-major order (Y Cb Cr per row) left late luma rows waiting for a CTA
// while chroma rows far ahead of their turn held them (2160p I picture: row 32 started 340 us late).
__host__ __device__ inline int intra_job_index(int row, int comp, int rows)
{
  const int H = (rows + 1) >> 1;
  if (comp == 0) return row + 2 * ((row + 1) >> 1);
  if (row < H) return 4 * row + comp;
  return rows + 2 * H + 2 * (row - H) + (comp - 1);
}
__host__ __device__ inline void intra_job_decode(int job, int rows, int& row, int& comp)
{
  const int H = (rows + 1) >> 1, head = rows + 2 * H;
  if (job < head)
  {
    const int q = job >> 2, m = job & 3;
    if (m == 0) { row = 2 * q; comp = 0; }
    else if (m == 3) { row = 2 * q + 1; comp = 0; }
    else { row = q; comp = m; }
  }
  else { const int t = job - head; row = H + (t >> 1); comp = 1 + (t & 1); }
}

struct PlaneSet
{
  int16_t* p[3];
  int      pitch[3];   // in samples
};

// Everything a kernel needs, passed by value as a __grid_constant__ kernel parameter (~1 KB).
struct FrameParams
{
  hmr_frame_hdr hdr;
  int w[3], h[3];          // component sizes in samples
  int csx, csy;            // chroma subsampling shifts
  int ctus_w, ctus_h;
  int w4, h4, w8;          // BS-map / QP-map strides
  PlaneSet work;           // picture under construction (before SAO)
  PlaneSet out;            // DPB slot the finished picture goes to
  PlaneSet dpb[HMR_MAX_SLOTS];
  const hmr_tu*              tu;
  const int16_t*             coef;
  int16_t*                   resid;     // compact residual buffer, same offsets as coef
  const hmr_intra*           intra;
  const hmr_ctu_intra_range* irange;
  const hmr_pu*              pu;
  const uint32_t*            pu_prefix;
  hmr_pu*                    mc_tiles;        // [n_mc_tiles] scratch: one record per 16x16-luma tile (k_mc.cu pre-pass)
  const hmr_ctu*             ctu;
  const uint8_t*             bs;
  const int8_t*              qp;
  const uint8_t*             cu_flags;
  const uint8_t*             scaling;         // scaling factors (hmr_records.h: HMR_SCALING_OFFSET) or nullptr
  const hmr_wp*              wp;              // explicit weighted prediction table [2][16][3] or nullptr
  const uint8_t*             pu_refidx;       // [n_pu] refIdx L0 | L1 << 4 (weighted prediction only)
  uint8_t*                   mc_tile_refidx;  // [n_mc_tiles] scratch: the same per tile
  uint4*                     intra_ops;       // [n_intra] decoded intra TUs           } scratch written by k_intra.cu's pre-pass
  uint16_t*                  intra_tab;       // [3][n_ctu][4352] reference-address tables }
  uint4*                     intra_prep;      // [3][n_ctu] residual span / table length   }
  int                        intra_max_rec;   // largest number of intra records of one (component, CTU) of this picture   } shared-memory
  int                        intra_max_addr;  // largest reference-address table of one (component, CTU), entries           } capacities of
  int                        intra_res_span;  // largest residual span of one (component, CTU), samples                      } intra_kernel
  unsigned long long*        intra_progress;  // [3][ctus_h], (epoch << 32) | CTUs finished in that row; [3 * ctus_h] = the job counter (zeroed by the pre-pass)
  uint32_t                   intra_job_mask[HMR_INTRA_JOB_WORDS];   // bit intra_job_index(row, comp, ctus_h): that CTU row of that component has intra TUs (a "job" of intra_kernel)
  int                        intra_jobs;      // number of set bits; -1 = every (row, component) is a job (mask not filled in: more rows than it has bits, or records not inspected)
  unsigned long long         epoch;
};

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(hi, max(lo, v)); }

// launchers (one per kernel file)
int  launch_mc(const FrameParams& P, cudaStream_t s);                 // returns number of launches
int  launch_resid(const FrameParams& P, cudaStream_t s);             // returns number of launches
cudaError_t launch_intra(const FrameParams& P, cudaStream_t s);
void launch_deblock(const FrameParams& P, int dir, cudaStream_t s);
void launch_sao(const FrameParams& P, cudaStream_t s);
void launch_hash(const PlaneSet& pic, const int w[3], const int h[3], const int bd[3], int type, uint32_t* d_out, uint32_t* d_scratch, cudaStream_t s);
int  intra_max_coresident_blocks(int device);
size_t intra_table_bytes(int nctu);
struct IntraSizes { int maxRec, maxAddr, resSpan; int jobs; uint32_t jobMask[HMR_INTRA_JOB_WORDS]; };
IntraSizes intra_sizes_host(const hmr_frame_hdr& h, const hmr_intra* rec, const hmr_ctu_intra_range* range);
size_t launch_pack(const PlaneSet& pic, const int w[3], const int h[3], int csx, int csy, int ncomp, const int bdInternal[3], const int bdOut[3],
                   const int crop[4], uint8_t* d_dst, cudaStream_t s);

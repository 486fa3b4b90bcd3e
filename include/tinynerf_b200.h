/*
 * tinynerf_b200 -- C ABI of the B200-native (sm_100a) packed-ray render/train hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  All pointers are DEVICE
 * pointers on the calling thread's current CUDA device unless marked [host].  Every entry point
 * returns 0 on success or a negative TNF_E_* code; tnf_last_error() gives a thread-local message.
 * Entry points are re-entrant, keep no global state, and launch on the caller's stream
 * (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream).
 *
 * Each function cites the reference interface (loicmagne/tinynerf, paths relative to the reference
 * root) it replaces.  Conventions shared by all of them:
 *   - floats are fp32, indices int32, "packing info" is int32 [n_rays][2] = (start, count)
 *   - packed samples are row-major [n][7] = (x,y,z contracted to [-1,1], dx,dy,dz, step)
 */
#ifndef TINYNERF_B200_H_
#define TINYNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNF_OK            0
#define TNF_E_INVALID    -1   /* bad argument (null pointer, negative size, bad enum) */
#define TNF_E_CUDA       -2   /* a CUDA runtime call / kernel launch failed */
#define TNF_E_UNSUPPORTED -3  /* device is not sm_100 or feature not built */

/* ---- library ------------------------------------------------------------------------------- */

/* ABI version (major*1000+minor). */
int tnf_version(void);
/* Thread-local message for the last non-zero return code on this thread ([host] string). */
const char* tnf_last_error(void);
/* Fills [host] ints: SM count, major, minor of the current device. */
int tnf_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* The persistent kernels launch one CTA per SM and split their tiles statically.  While a collective's kernel occupies
 * some SMs (a gradient all-reduce running under the weight-gradient kernels), CTAs that find no free SM start a whole wave
 * late; capping the grid at the SMs actually free avoids that.  n_sms <= 0 removes the cap.  Per calling thread; returns
 * the previous value. */
int tnf_set_sm_budget(int n_sms);
/* Diagnostics: select an alternative kernel for the calling thread (which: 0 = weight gradient with both operands in
 * shared memory, 1 = one-thread-per-texel TV kernel; value 1 = on, 0 = default).  Returns the previous value, -1 for an
 * unknown `which`.  The library never reads the environment. */
int tnf_set_variant(int which, int value);

/* ---- a1/a2: NeRF-equation weights over packed rays -------------------------------------------
 * Replaces src/cuda.cu:66-95 (compute_weights_fwd, kernel :3-30) and src/cuda.cu:97-132
 * (compute_weights_bwd, kernel :32-58).
 *
 *   weights[k] = T_k * (1 - a_k),  a_k = __expf(-sigmas[k]*steps[k]),  T_k = prod_{j<k in ray} a_j
 *   a ray stops at the first k with T_k <= threshold; every later sample of that ray gets 0.
 *   grad_sigmas[k] = steps[k] * (T_{k+1} * g_k - sum_{j>k in ray} weights[j]*g_j)   (no termination)
 *
 * `steps` may be strided (steps_stride in elements; 7 for the step column of packed samples, which
 * removes the reference's .contiguous() copy at src/core.py:196).  Outputs are caller-allocated and
 * EVERY element is written (no memset pass needed).
 *
 * flags:
 *   TNF_W_TRUSTED_PARTITION  caller guarantees info is a sorted, gap-free partition of [0,n_samples)
 *                            (what RayProvider emits); skips validation and the fallback launch.
 *   TNF_W_NO_EXACT_TERMINATION  skip the serial re-evaluation of rays whose transmittance comes
 *                            within rounding distance of `threshold` (with it ON the
 *                            (weights>0) mask is bit-identical to the reference's serial kernel).
 * `status` is a 4-byte device word used as scratch (may be NULL iff TNF_W_TRUSTED_PARTITION);
 *   after the call bit0 is set when info was NOT a partition and the generic ray-serial kernel
 *   produced the result instead.
 */
#define TNF_W_TRUSTED_PARTITION     1
#define TNF_W_NO_EXACT_TERMINATION  2

int tnf_weights_fwd(const float* sigmas, const float* steps, int64_t steps_stride,
                    const int32_t* info, float threshold, float* weights,
                    int64_t n_samples, int64_t n_rays, int flags, uint32_t* status, void* stream);

int tnf_weights_bwd(const float* sigmas, const float* steps, int64_t steps_stride,
                    const int32_t* info, const float* weights, const float* grad_weights,
                    float* grad_sigmas, int64_t n_samples, int64_t n_rays, int flags,
                    uint32_t* status, void* stream);

/* ---- a4-a8,a10: ray marching + contraction + occupancy lookup + sample packing ----------------
 * Replaces RayProvider.__call__ (src/core.py:165-188) together with RayMarcherAABB.__call__
 * (:73-88), RayMarcherUnbounded.__call__ (:48-59), ContractionAABB (:27-31), ContractionMip360
 * (:16-20, order=inf) and OccupancyGrid.forward (:148-156).
 *
 * Two calls per ray batch.  tnf_march_count evaluates the [n_rays, n_steps] sample lattice, writes
 * the occupancy&bounds mask as a bitfield (n_rays x ceil(n_steps/32) words) and the packing info
 * (exclusive int32 scan of per-ray counts + info_offset -- the `info[:,0] += current_size` of the
 * dynamic-batch accumulator, src/run.py:231) and the total in *n_packed (device int64).  The caller
 * reads n_packed, allocates packed[n_packed][7], and tnf_march_pack fills it in (ray, step) order.
 */
#define TNF_SCENE_AABB       0
#define TNF_SCENE_UNBOUNDED  1

typedef struct tnf_march_params {
  int32_t scene;            /* TNF_SCENE_* */
  int32_t n_steps;          /* samples per ray S */
  float   aabb[6];          /* AABB: min xyz, max xyz (scene==AABB) */
  float   near, far;        /* AABB: clamp of t_min (src/core.py:81) */
  float   step_size;        /* AABB: ||aabb1-aabb0||/S as computed by the host (src/core.py:68-70) */
  const float* t_table;     /* UNBOUNDED: device [S] t values   (src/core.py:52-57, ray independent) */
  const float* step_table;  /* UNBOUNDED: device [S] step sizes */
  const float* grid;        /* occupancy grid, device fp32 [gd][gh][gw] */
  int32_t gd, gh, gw;
  float   threshold;        /* min(base_threshold, mean) as fp32 (src/core.py:127,156) */
  const float* noise;       /* optional device [n_rays][S] U[0,1) jitter (training); NULL = see seed */
  int32_t jitter;           /* 0: no jitter (training=False); 1: jitter from `noise` or Philox(seed) */
  uint64_t seed, offset;    /* Philox4x32-10 key/counter base when jitter && !noise */
  const float* threshold_dev; /* optional DEVICE scalar that replaces `threshold` (e.g. min(base, grid.mean()) computed on the
                               device right after an occupancy update, so the update needs no host synchronisation) */
  const uint32_t* empty_bits; /* optional DEVICE bitfield from tnf_occ_build_empty_bits for THIS grid and threshold (NULL: none).
                               tnf_march_count stages it in shared memory and skips the float lookup for lattice points in
                               blocks it marks certainly empty; the keep-mask is bit-identical with and without it. */
} tnf_march_params;

/* "Certainly empty" classifier of an occupancy grid (the mask of src/core.py:148-156 is trilinear(grid) > threshold on float
 * values, so a bit test can only short-circuit where the outcome is certain).  bits: tnf_occ_empty_bits_words(gd,gh,gw) 32-bit
 * words, two levels back to back: coarse [ceil(gd/2)][ceil(gh/2)][ceil(ceil(gw/2)/32)] -- bit (bz,by,bx) set when every lattice
 * value a lookup in the 2x2x2-cell block can touch is <= threshold*(1 - 1e-5) (32 KB for 128^3, staged in shared memory by
 * tnf_march_count) -- then fine [gd][gh][ceil(gw/32)], one bit per cell (its 8 corners).  Where a bit is set the lookup's convex
 * combination cannot exceed the threshold in fp32 either.  Rebuild after every change of the grid or of the threshold
 * (threshold_dev, when given, is read by the kernel). */
int64_t tnf_occ_empty_bits_words(int32_t gd, int32_t gh, int32_t gw);
int tnf_occ_build_empty_bits(const float* grid, int32_t gd, int32_t gh, int32_t gw, float threshold,
                             const float* threshold_dev /*optional*/, uint32_t* bits, void* stream);

int tnf_march_count(const tnf_march_params* p /*[host]*/, const float* rays_o, const float* rays_d,
                    int64_t n_rays, int32_t info_offset, uint32_t* mask_bits, int32_t* info,
                    int64_t* n_packed, void* stream);

int tnf_march_pack(const tnf_march_params* p /*[host]*/, const float* rays_o, const float* rays_d,
                   int64_t n_rays, int32_t info_offset, const uint32_t* mask_bits,
                   const int32_t* info, float* packed, float* steps_out /*optional [n]*/,
                   int32_t* ray_idx_out /*optional [n]*/, int64_t n_packed, void* stream);

/* OccupancyGrid.forward alone (src/core.py:148-156): out[i] = trilinear(grid, coords[i]) > thr.
 * coords are [-1,1] with coords[:,0] -> W (last grid dim).  Optionally returns the interpolated
 * value (values may be NULL). */
int tnf_occ_query(const float* grid, int32_t gd, int32_t gh, int32_t gw, const float* coords,
                  int64_t n, float threshold, uint8_t* out_mask, float* values, void* stream);

/* ---- a9: occupancy grid update / decay --------------------------------------------------------
 * Replaces the body of OccupancyGrid.update (src/core.py:134-145), split around the sigma_fn call:
 *   tnf_occ_update_coords: coords[c] = -1 + 2*(cell_xyz[c] + noise[c]) / size       (:137)
 *   tnf_occ_update_apply : alpha = 1 - exp(-sigma*step); grid = alpha>thr ? 1 : decay*grid  (:139-144)
 * operating on cells [cell0, cell0+n_cells) of the flattened [gd][gh][gw] grid.  `size3` is the
 * reference's self.size = (gd,gh,gw) that divides (x,y,z) (quirk kept, src/core.py:109,137).
 * noise: device [n_cells][3] U[0,1) or NULL for Philox(seed, offset).
 */
int tnf_occ_update_coords(int32_t gd, int32_t gh, int32_t gw, int64_t cell0, int64_t n_cells,
                          const float* noise, uint64_t seed, uint64_t offset, float* coords,
                          void* stream);
int tnf_occ_update_apply(float* grid, int64_t cell0, int64_t n_cells, const float* sigma,
                         float step_size, float threshold, float decay, void* stream);
/* same, with the threshold optionally read from a device scalar (the value min(base, mean) left on the device by the
 * previous update: a training loop then never synchronises the host for the grid update) */
int tnf_occ_update_apply_dev(float* grid, int64_t cell0, int64_t n_cells, const float* sigma, float step_size,
                             float threshold, const float* threshold_dev /*optional*/, float decay, void* stream);

/* ---- a12: K-Planes fused feature lookup -------------------------------------------------------
 * Replaces KPlanesFeatureField.forward (src/models.py:153-163) = 9x KPlanesFeaturePlane.forward
 * (:105-113, F.grid_sample bilinear/zeros/align_corners=True) + 3 Hadamard products + concat, and
 * its autograd backward (scatter-add of plane gradients).
 *
 * planes[s*3+p] points to plane p of scale s stored CHANNELS-LAST: [res_s][res_s][C] fp32 (the
 * logical parameter stays [1,C,res,res]; see DESIGN.md "data layout").  Plane p uses coordinate
 * pair (0,1),(0,2),(1,2): first -> W/x axis, second -> H/y axis.  C must be a multiple of 4, <=32.
 * x is [n][x_stride] with xyz in the first 3 columns (x_stride=7 reads packed samples in place).
 * out / grad_out are [n][n_scales*C].  Backward ACCUMULATES into grad_planes (same layout).
 */
int tnf_kplanes_fwd(const float* const* planes /*[host] n_scales*3 device ptrs*/,
                    const int32_t* res /*[host] n_scales*/, int32_t n_scales, int32_t channels,
                    const float* x, int64_t x_stride, int64_t n, float* out, void* stream);
int tnf_kplanes_bwd(const float* const* planes, float* const* grad_planes, const int32_t* res,
                    int32_t n_scales, int32_t channels, const float* x, int64_t x_stride, int64_t n,
                    const float* grad_out, void* stream);

/* tnf_kplanes_bwd restricted to scales [scale_begin, scale_end): lets a data-parallel caller start the all-reduce of one
 * scale's plane gradients while the other scales are still being scattered (grad_out keeps its full [n][n_scales*C] rows). */
int tnf_kplanes_bwd_scales(const float* const* planes, float* const* grad_planes, const int32_t* res, int32_t n_scales,
                           int32_t channels, const float* x, int64_t x_stride, int64_t n, const float* grad_out,
                           int32_t scale_begin, int32_t scale_end, void* stream);

/* Diagnostics (scripts/time_kplanes.py): partial / alternative forms of tnf_kplanes_bwd used to
 * measure what bounds it.  mode 1 = gather + blend only (no reductions), 2 = red.global.add.v4.f32 reductions only (no plane
 * reads), 3 = full backward with the reductions issued by the TMA engine (rows staged in shared memory, one
 * cp.reduce.async.bulk add.f32 per pair of x-adjacent corners).  Mode 3 computes the same gradients as tnf_kplanes_bwd. */
int tnf_kplanes_bwd_ex(const float* const* planes, float* const* grad_planes, const int32_t* res, int32_t n_scales,
                       int32_t channels, const float* x, int64_t x_stride, int64_t n, const float* grad_out, int32_t mode,
                       void* stream);

/* ---- stand-alone forms of the reference's public helper callables (csrc/helpers.cu) ----------------------------
 * Not on the training hot path (RayProvider / the feature fields fuse these steps); they back the methods of the host-side
 * mirror that callers may use on their own, so that no method of the mirror runs PyTorch arithmetic.
 *   tnf_marcher_aabb      RayMarcherAABB.__call__ (src/core.py:73-88): t_values, step_sizes [n_rays][n_steps]
 *   tnf_contract          ContractionAABB.__call__ (src/core.py:27-31; mask [n] bytes) / ContractionMip360.__call__
 *                         with order = inf (src/core.py:16-20; mask must be NULL)
 *   tnf_plane_lookup_*    KPlanesFeaturePlane.forward (src/models.py:105-113): one channels-last [h][w][C] plane,
 *                         xy [n][xy_stride] (x -> w axis, y -> h axis) -> out [n][C]; bwd ACCUMULATES into grad_plane
 *   tnf_grid3_lookup_*    CobafaGrid.forward (src/models.py:228-238): one channels-last [d][h][w][C] grid
 *   tnf_abs_mean_*        KPlanesFeaturePlane.loss_l1 (src/models.py:120-121): *sum = sum |x| (double; the caller divides
 *                         by n); bwd writes grad_x = sign(x) * (*grad_scale) / n */
int tnf_marcher_aabb(const float* aabb6 /*[host] min xyz, max xyz*/, float near_, float far_, float step_size,
                     const float* rays_o, const float* rays_d, int64_t n_rays, int32_t n_steps, float* t_values,
                     float* step_sizes, void* stream);
int tnf_contract(int32_t scene, const float* aabb6 /*[host], AABB scenes*/, const float* coords, int64_t n, float* out,
                 uint8_t* mask, void* stream);
int tnf_plane_lookup_fwd(const float* plane, int32_t h, int32_t w, int32_t channels, const float* xy, int64_t xy_stride,
                         int64_t n, float* out, void* stream);
int tnf_plane_lookup_bwd(float* grad_plane, int32_t h, int32_t w, int32_t channels, const float* xy, int64_t xy_stride,
                         int64_t n, const float* grad_out, void* stream);
int tnf_grid3_lookup_fwd(const float* grid, int32_t d, int32_t h, int32_t w, int32_t channels, const float* x,
                         int64_t x_stride, int64_t n, float* out, void* stream);
int tnf_grid3_lookup_bwd(float* grad_grid, int32_t d, int32_t h, int32_t w, int32_t channels, const float* x,
                         int64_t x_stride, int64_t n, const float* grad_out, void* stream);
/* PositionalEncoding.forward (src/models.py:30-39): x [n][x_stride] (first d columns) -> out [n][ld_out], columns
 * c*2*n_freqs + k = sin(2^k pi x_c), c*2*n_freqs + n_freqs + k = cos(2^k pi x_c). */
int tnf_positional_encoding(const float* x, int64_t x_stride, int32_t d, int32_t n_freqs, int64_t n, float* out,
                            int64_t ld_out, void* stream);
int tnf_abs_mean_fwd(const float* x, int64_t n, double* sum, void* stream);
int tnf_abs_mean_bwd(const float* x, int64_t n, const float* grad_scale, float* grad_x, void* stream);

/* ---- a13: K-Planes total-variation regulariser ---------------------------------------------------
 * Replaces KPlanesFeaturePlane.loss_tv / KPlanesFeatureField.loss_tv (src/models.py:115-118,165-172)
 * and their autograd backward for a table of n_planes channels-last planes [res][res][C].
 * fwd : sums[2*i+0] = sum (p[h+1,w]-p[h,w])^2, sums[2*i+1] = sum (p[h,w+1]-p[h,w])^2  (device doubles;
 *       loss_tv(plane i) = sums[2i]/(C*(res-1)*res) + sums[2i+1]/(C*res*(res-1)))
 * bwd : grads[i] (=|+=) plane_weight[i] * (*gscale) * d loss_tv(plane i)/d plane; gscale is a DEVICE
 *       scalar (the upstream gradient), plane_weight a [host] array (NULL = 1), e.g. 1/9 for the field mean.
 */
int tnf_tv_fwd(const float* const* planes /*[host]*/, const int32_t* res /*[host]*/, int32_t n_planes,
               int32_t channels, double* sums, void* stream);
int tnf_tv_bwd(const float* const* planes, float* const* grads, const int32_t* res, int32_t n_planes,
               int32_t channels, const float* plane_weight /*[host]*/, const float* gscale, int32_t accumulate,
               void* stream);
/* tnf_tv_bwd that also returns tnf_tv_fwd's `sums` from the same pass over the planes (loss value for reporting +
 * gradient: one read of the planes instead of two). */
int tnf_tv_fwd_bwd(const float* const* planes, float* const* grads, const int32_t* res, int32_t n_planes,
                   int32_t channels, const float* plane_weight /*[host]*/, const float* gscale, int32_t accumulate,
                   double* sums, void* stream);

/* ---- Adam step over a table of tensors ------------------------------------------
 * Replaces torch.optim.Adam.step as configured at src/run.py:186 (L2 weight decay folded into the
 * gradient, bias correction, no amsgrad) for n_tensors flat fp32 tensors in one launch:
 *   g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * All tables are [host] arrays of device pointers; `step` is the 1-based step count t.
 */
int tnf_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int64_t* numel, int32_t n_tensors, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, void* stream);
/* Same update with the grid capped at max_blocks 256-thread blocks that stride over the work (0 = one block per 4096
 * elements): a small persistent grid (one block per SM) can share the SMs with a one-CTA-per-SM tensor-core kernel on another
 * stream instead of queueing in front of its CTAs. */
int tnf_adam_step_grid(float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* numel, int32_t n_tensors, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int64_t step, int32_t max_blocks, void* stream);

/* Sorted scatter for the scales below the finest (same result as tnf_kplanes_bwd up to fp32 summation order).
 * tnf_kplanes_sort: for each orientation o of itertools.combinations(range(3), 2) a counting sort of the n samples by their
 * cell at resolution sort_res -> pos [3][n] int32 (slot of sample i in orientation o's order) and uv [3][n][2] (the
 * orientation's two coordinates in slot order).  scratch: tnf_kplanes_sort_scratch_ints(sort_res, n) int32 words.
 * tnf_kplanes_bwd_sorted: scales [0, sorted_scales) go through `rows` ([3*sorted_scales][n][32] floats: their per-plane
 * gradient rows in slot order) and are reduced run by run -- one red.v4 per corner per run of samples in the same cell
 * instead of per sample; the other scales scatter directly.  phase 0 = both kernels, 1 = gather + direct scatters + row
 * stores only, 2 = the run-merging scatter of the stored rows only (callers that start work between the two).  32 channels. */
int64_t tnf_kplanes_sort_scratch_ints(int32_t sort_res, int64_t n);
int tnf_kplanes_sort(const float* x, int64_t x_stride, int64_t n, int32_t sort_res, int32_t* scratch, int32_t* pos, float* uv,
                     void* stream);
int tnf_kplanes_bwd_sorted(const float* const* planes, float* const* grad_planes, const int32_t* res, int32_t n_scales,
                           int32_t channels, const float* x, int64_t x_stride, int64_t n, const float* grad_out,
                           int32_t sorted_scales, const int32_t* pos, const float* uv, float* rows, int32_t phase, void* stream);

/* ---- a14: Cobafa fused basis/coefficient lookup ----------------------------------------------
 * Replaces CobafaFeatureField.forward up to the concat (src/models.py:258-264): coef = trilinear
 * (coef_grid, x); y_l = trilinear(basis_l, 2*((f_l*x) mod 1)-1) * coef[l]; out = cat_l y_l.
 * Grids are CHANNELS-LAST [r][r][r][c] fp32.  Backward accumulates into grad_basis/grad_coef.
 */
int tnf_cobafa_fwd(const float* const* basis /*[host] L device ptrs*/, const int32_t* basis_res,
                   const int32_t* basis_ch, const float* freqs /*[host] L*/, int32_t n_levels,
                   const float* coef, int32_t coef_res, const float* x, int64_t x_stride, int64_t n,
                   float* out, void* stream);
int tnf_cobafa_bwd(const float* const* basis, float* const* grad_basis, const int32_t* basis_res,
                   const int32_t* basis_ch, const float* freqs, int32_t n_levels, const float* coef,
                   float* grad_coef, int32_t coef_res, const float* x, int64_t x_stride, int64_t n,
                   const float* grad_out, void* stream);

/* ---- a15-a17: dense layers of the MLP heads on tcgen05 tensor cores (3xTF32, fp32 accumulate in TMEM) ----
 * Replace the nn.Linear (+ReLU) stacks of MLP / VanillaOpacityDecoder / VanillaColorDecoder / the Cobafa trunk
 * (src/models.py:7-28,70-89,255,266) and their autograd backward.  Row-major fp32, weight = nn.Linear.weight
 * [n, k]; k <= 160, n in {32,64,96,128}; x/y/dy/dx 16-byte aligned with leading dimensions multiple of 4.
 *   fwd  : y = act(x W^T + b) (relu flag); optionally a fused small head on the (activated) y:
 *          head_out[m, o] = head_act(sum_j y[m,j] head_w[o,j] + head_b[o]), n_head <= 4,
 *          head_act 0 = identity, 1 = exp(v - 1) (truncated_exp(x - 1.), src/models.py:74), 2 = sigmoid (:85)
 *   dgrad: dx = (dy W) (* (relu_src > 0) when relu_src != NULL: ReLU backward of the producer of x)
 *   wgrad: dweight += dy^T x ; dbias += column sums of dy      (atomic accumulation: zero them first)
 *   head_bwd: gradients of the fused head: dh (masked by h > 0), dhead_w, dhead_b (accumulated)
 * Layers with n == 128 and k <= 128 (the Cobafa trunk) run with the weights stationary in tensor memory (csrc/wstat.cu: the
 * transposed GEMM, W as the TMEM A operand, sample tiles and the ReLU mask by TMA) and their weight gradient in one pass over
 * dy and x (wgrad128_tma_kernel); same arithmetic, same arguments -- the entry points choose by shape, never by backend.
 */
/* Layers wider than the resident-weight kernels above cover (in_features > 160 or out_features > 128: the reference's
 * VanillaFeatureMLP(10, 256, 8), src/models.py:59-68, the decoders behind it and the Cobafa colour head's 179-wide first
 * layer): the same 3xTF32 tcgen05 arithmetic with both operands streamed (csrc/wide.cu).  Any m, n, k >= 1; row-major fp32;
 * weight [n][ldw]; act 0 = none, 1 = ReLU, 2 = exp(v - 1), 3 = sigmoid; dgrad masks with (relu_src > 0) when given;
 * wgrad ACCUMULATES into dweight [n][lddw] / dbias [n] (zero them first). */
int tnf_wide_linear_fwd(const float* x, int64_t ldx, const float* weight, int64_t ldw, const float* bias, float* y, int64_t ldy,
                        int64_t m, int32_t n, int32_t k, int32_t act, void* stream);
int tnf_wide_linear_bwd_data(const float* dy, int64_t lddy, const float* weight, int64_t ldw, float* dx, int64_t lddx,
                             const float* relu_src, int64_t ldrs, int64_t m, int32_t n, int32_t k, void* stream);
int tnf_wide_linear_bwd_weight(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dweight, int64_t lddw,
                               float* dbias, int64_t m, int32_t n, int32_t k, void* stream);
int tnf_linear_fwd(const float* x, int64_t ldx, const float* weight, const float* bias, float* y, int64_t ldy,
                   int64_t m, int32_t n, int32_t k, int32_t relu, const float* head_w, const float* head_b,
                   float* head_out, int32_t n_head, int32_t head_act, void* stream);
int tnf_linear_bwd_data(const float* dy, int64_t lddy, const float* weight, float* dx, int64_t lddx,
                        const float* relu_src, int64_t ldrs, int64_t m, int32_t n, int32_t k, void* stream);
int tnf_linear_bwd_weight(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dweight, float* dbias,
                          int64_t m, int32_t n, int32_t k, void* stream);
/* Several 64-output weight gradients over the SAME m rows in one launch (the heads' dW_i = dh_i^T h_{i-1}, src/models.py:7-28
 * backward): job j adds dy[j]^T x[j] into dweight[j] [64, k[j]] and the column sums of dy[j] into dbias[j] (dbias or dbias[j]
 * NULL to skip).  All tables are [host] arrays of n_jobs <= 4 entries; dy[j] [m,64], x[j] [m,k[j]] with k[j] <= 128, 16-byte
 * aligned, leading dimensions multiple of 4.  Same arithmetic as n_jobs calls of tnf_linear_bwd_weight with n = 64; one cold
 * start and one drain instead of n_jobs. */
int tnf_linear_bwd_weight_multi(int32_t n_jobs, const float* const* dy /*[host]*/, const int64_t* lddy /*[host]*/,
                                const float* const* x /*[host]*/, const int64_t* ldx /*[host]*/, const int32_t* k /*[host]*/,
                                float* const* dweight /*[host]*/, float* const* dbias /*[host], optional*/, int64_t m, void* stream);
/* wgrad of a layer whose input is a concatenation that was never materialised: x = [xa (ka columns) | xb (kb columns)],
 * dweight [n, ka+kb] += dy^T x, dbias += column sums of dy (NULL to skip).  The colour head's first layer reads
 * cat([PE(d), d], features) (src/models.py:87): xa = the [PE(d) | d] rows, xb = the feature rows.  n must be 64.
 * scratch: optional device buffer of tnf_wgrad_cat_scratch_bytes(ka, kb) bytes, 16-byte aligned, ZERO before the first
 * call (every call leaves it zero again).  With it the per-CTA partial sums are reduced with 16-byte operations even
 * though rows of dweight (ka+kb floats) and the column ka are not 16-byte aligned; without it they are scalar atomics. */
int64_t tnf_wgrad_cat_scratch_bytes(int32_t ka, int32_t kb);
int tnf_linear_bwd_weight_cat(const float* dy, int64_t lddy, const float* xa, int64_t ldxa, int32_t ka, const float* xb,
                              int64_t ldxb, int32_t kb, float* dweight, float* dbias, int64_t m, int32_t n, float* scratch,
                              void* stream);
/* Input row of VanillaColorDecoder.forward (src/models.py:87): out[m] = [PE_{n_freqs}(dirs[m]) | dirs[m] | feats[m]],
 * zero-padded to ld_out floats (PositionalEncoding layout, src/models.py:36-39). */
int tnf_color_input(const float* dirs, int64_t ld_dirs, const float* feats, int64_t ld_feats, int32_t n_freqs,
                    int32_t feat_dim, float* out, int64_t ld_out, int64_t n, void* stream);
int tnf_head_bwd(const float* h, int64_t ldh, const float* head_w, const float* out, const float* dout, float* dh,
                 float* dhead_w, float* dhead_b, int64_t m, int32_t n, int32_t n_head, int32_t head_act, void* stream);

/* Both decoder heads of the K-Planes / Cobafa pipelines in one persistent kernel (forward):
 *   sigma = exp(W_s1 relu(W_s0 f + b_s0) + b_s1 - 1)                   VanillaOpacityDecoder.forward, src/models.py:76-77
 *   rgb   = sigmoid(W_c4 relu(W_c3 relu(W_c2 relu(W_c1 relu(W_c0 x + b_c0) ...))))   VanillaColorDecoder.forward, :86-89
 * feats [m, feat_dim] are the feature rows f, xc [m, k0] the colour-input rows x = [PE(d) | d | f] (tnf_color_input).
 * color_w/color_b: [host] arrays of 5 device pointers (W_c0 [64,k0], W_c1..3 [64,64], W_c4 [3,64] and biases);
 * sigma_w/sigma_b: 2 device pointers (W_s0 [64,feat_dim], W_s1 [1,64]).  Hidden width is 64, three hidden colour layers
 * (the reference's VanillaColorDecoder(8, dim, 64, 3) / VanillaOpacityDecoder(dim), src/run.py:131-150).
 * h_out: optional [host] array of 4 device pointers [m,64] receiving the colour head's hidden activations, hs_out [m,64]
 * the density head's (the backward kernels read them); rgb [m,3], sigma [m].  workspace: device scratch of
 * tnf_heads_workspace_bytes(feat_dim, k0) bytes (packed weight images), 16-byte aligned.
 * xc_cols: how many leading columns of the colour-input row `xc` holds.  xc_cols == k0: the whole row (tnf_color_input with
 * the features).  xc_cols == k0 - feat_dim: only [PE(d) | d]; the trailing feat_dim columns of the row are the feature row
 * itself and are taken from `feats` (the concatenation is never written). */
int64_t tnf_heads_workspace_bytes(int32_t feat_dim, int32_t k0);
int tnf_heads_fwd(const float* feats, int64_t ld_feats, int32_t feat_dim, const float* xc, int64_t ld_xc, int32_t k0,
                  int32_t xc_cols, const float* const* color_w /*[host]*/, const float* const* color_b /*[host]*/,
                  const float* const* sigma_w /*[host]*/, const float* const* sigma_b /*[host]*/,
                  float* const* h_out /*[host], optional*/, float* hs_out /*optional*/, float* rgb, float* sigma, int64_t m,
                  void* workspace, void* stream);

/* The data-gradient ("dgrad") chain of both heads in one persistent kernel: from the gradients at the last hidden layers
 * (dh3 [m,64] of the colour head, dhs [m,64] of the density head, as produced by tnf_head_bwd) to the gradient of the
 * feature rows that feed both heads:
 *   dh2 = (h2>0)*(dh3 W_c3), dh1 = (h1>0)*(dh2 W_c2), dh0 = (h0>0)*(dh1 W_c1),
 *   dfeat = dh0 W_c0[:, feat_col0 : feat_col0+feat_dim] + dhs W_s0
 * masks: [host] {h2, h1, h0} saved activations [m,64]; color_w: [host] {W_c0 [64,k0], W_c1, W_c2, W_c3 [64,64]};
 * dh_out: optional [host] {dh2, dh1, dh0} [m,64] (inputs of the weight-gradient kernels); dfeat [m, ld_dfeat].
 * Replaces five tnf_linear_bwd_data launches and the add of the two feature-gradient terms.  workspace:
 * tnf_heads_bwd_workspace_bytes(feat_dim) bytes of device scratch, 16-byte aligned. */
int64_t tnf_heads_bwd_workspace_bytes(int32_t feat_dim);
int tnf_heads_bwd_data(const float* dh3, const float* dhs, const float* const* masks /*[host]*/,
                       const float* const* color_w /*[host]*/, int32_t k0, int32_t feat_col0, const float* sigma_w0,
                       int32_t feat_dim, float* const* dh_out /*[host], optional*/, float* dfeat, int64_t ld_dfeat, int64_t m,
                       void* workspace, void* stream);

/* ---- a18: compositing (segment sums over packed rays) -----------------------------------------
 * Replaces the index_add_ block of NerfRenderer.forward (src/core.py:256-265; the reference's own
 * "TODO: cuda kernel this"):  rgb_ray = sum_k w_k*rgb_k ; opacity = sum_k w_k ;
 * out = rgb_ray + bg*(1-opacity) when bg != NULL ([host] 3 floats).
 * Colours are defined only where w_k > 0 (the reference evaluates the colour head on that subset and leaves 0
 * elsewhere, src/core.py:243-250); rgbs[k] is ignored where w_k <= 0, so callers may pass a dense evaluation.
 * Backward: grad_rgb[k] = w_k*go[ray] ; grad_w[k] = [w_k > 0] <rgb_k, go[ray]> - <bg, go[ray]>.
 */
int tnf_composite_fwd(const float* weights, const float* rgbs, const int32_t* info, int64_t n_samples,
                      int64_t n_rays, const float* bg, float* out_rgb, float* out_opacity, void* stream);
int tnf_composite_bwd(const float* weights, const float* rgbs, const int32_t* info, int64_t n_samples,
                      int64_t n_rays, const float* bg, const float* grad_out, float* grad_weights,
                      float* grad_rgbs, void* stream);

/* Loss of the training iteration and its gradient wrt the rendered colours (src/run.py:252 nn.MSELoss, :259
 * `scaler.scale(loss).backward()` with the never-unscaled factor `grad_scale`), normalised by the ray count of the
 * UNION batch of all ranks: denom = 3 * n_rays_global (read from the device scalar when given).
 *   *loss_out = sum (rendered - target)^2 / denom ;  grad_rendered = grad_scale * 2 (rendered - target) / denom */
int tnf_mse_loss_grad(const float* rendered, const float* target, int64_t n_rays, float n_rays_global,
                      const float* n_rays_global_dev /*optional*/, float grad_scale, float* grad_rendered /*optional*/,
                      float* loss_out /*optional*/, void* stream);

/* tnf_composite_fwd + tnf_mse_loss_grad + tnf_composite_bwd of one training iteration in a single pass over the rays (the
 * loss gradient of a ray depends only on that ray's rendered colour): out_rgb [n_rays,3], grad_weights [n], grad_rgbs [n,3],
 * *loss_out as tnf_mse_loss_grad, plus sum_i extra_terms[i]*extra_coef[i] (n_extra device doubles each; e.g. the TV sums of
 * tnf_tv_fwd_bwd and their weights, so the reported loss is complete without further kernels).  scratch: 16 bytes of device
 * memory, 8-byte aligned, ZERO before the first call (left zero by every call). */
int tnf_composite_loss_fwd_bwd(const float* weights, const float* rgbs, const int32_t* info, int64_t n_samples,
                               int64_t n_rays, const float* bg, const float* target, float n_rays_global,
                               const float* n_rays_global_dev /*optional*/, float grad_scale, float* out_rgb,
                               float* grad_weights, float* grad_rgbs, float* loss_out, void* scratch,
                               const double* extra_terms /*optional*/, const double* extra_coef /*optional*/, int32_t n_extra,
                               void* stream);

/* ---- (e) + 8f rank 1: the data-parallel parameter path over NVLink peer memory ----------------------------------------
 * New work (the reference trains on one device, src/run.py:98); the update it distributes is optimizer.step() of
 * torch.optim.Adam as configured at src/run.py:186, applied at src/run.py:258-261 to the gradients of the union batch.
 *
 * Every rank holds a flat fp32 gradient buffer and a flat fp32 parameter buffer at the SAME offsets in symmetric memory:
 * peer_grads[r] / peer_params[r] / peer_flags[r] are this process's mappings of rank r's buffers ([host] arrays of `world`
 * device pointers, entry `rank` being the local buffer); mc_grad / mc_param are the NVSwitch multicast addresses of the two
 * buffers, or both NULL (then the kernel loads from / stores to the peers one by one).  Called by every rank in the same
 * step on its own slice [lo, hi) (multiples of 4 elements; slices of different ranks must not overlap), it reduces the
 * slice's gradients over the ranks, applies tnf_adam_step's update to the local parameters / exp_avg / exp_avg_sq (flat,
 * indexed like the parameters; only the owning rank's slice is ever touched) and writes the new parameters to every rank --
 * reduce-scatter + sharded Adam + all-gather in one pass, bracketed by two rank barriers (flags: [n_ctas][TNF_DP_MAX_RANKS]
 * uint32 words per rank, n_ctas <= 4 x SM count, zero before the first call; each call consumes the epochs `epoch` and `epoch + 1`, so successive
 * calls on one flag pad pass epoch = 1, 3, 5, ...).  n_ctas identical on every rank.  *error (local device
 * int32) becomes non-zero if a barrier gave up after ~4 s (a peer died): the result is then undefined. */
#define TNF_DP_MAX_RANKS   16
#define TNF_DP_COUNT_SLOTS 4
int tnf_dp_reduce_adam_bcast(const float* const* peer_grads, float* const* peer_params, const float* mc_grad,
                             float* mc_param, float* exp_avg, float* exp_avg_sq, int64_t lo, int64_t hi,
                             uint32_t* const* peer_flags, int32_t n_ctas, int32_t rank, int32_t world, uint32_t epoch,
                             int32_t* error, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                             void* stream);
/* Ray count of the union batch (the MSE normaliser of src/run.py:252 when ranks hold different ray counts).  peer_slots[r]:
 * this process's mapping of rank r's slot table, uint64 [TNF_DP_COUNT_SLOTS][TNF_DP_MAX_RANKS] in symmetric memory, zero
 * before the first call.  publish: stores {step, count} into entry [slot][rank] of every rank's table (step != 0; use
 * slot = step % TNF_DP_COUNT_SLOTS).  sum: waits until the local table holds every rank's word of `step`, then writes the
 * total to *out (float, exact for counts < 2^24); *error = 2 on a ~4 s timeout. */
int tnf_dp_publish_count(uint64_t* const* peer_slots, int32_t world, int32_t rank, int32_t slot, uint32_t step, float count,
                         void* stream);
int tnf_dp_sum_counts(const uint64_t* slots, int32_t world, int32_t slot, uint32_t step, float* out, int32_t* error,
                      void* stream);

/* ---- ray batches (DataLoader collate + .to(device), src/run.py:116-122,226-228) --------------------------------------------
 * out[i][:] = table[idx[i]][:] for n_rows rows of row_floats floats.  idx: DEVICE int64 (entries in [0, n_table_rows), not
 * checked).  `table` is a device pointer OR a pointer into pinned host memory (cudaHostAlloc / torch pin_memory under unified
 * addressing): the kernel then reads the rows over the host link (zero-copy) -- the H2D transfer of exactly these rows. */
int tnf_gather_rows(const float* table, int64_t n_table_rows, int32_t row_floats, const int64_t* idx, int64_t n_rows, float* out,
                    void* stream);

/* ---- host helper: lazily shuffled ray order (DataLoader(shuffle=True), src/run.py:116-122) --------------------------------
 * HOST pointers.  perm: n entries, initially 0..n-1 (any permutation); out receives `count` ray indices.  pos is a global
 * position counter (a multiple of world; advance it by count*world), *fresh_from the first position never drawn so far
 * (start at 0), rng_state one 64-bit word of generator state.  Positions below *fresh_from are replayed unchanged. */
int tnf_shuffle_next(int64_t* perm, int64_t n, int64_t pos, int64_t count, int32_t rank, int32_t world, int64_t* fresh_from,
                     uint64_t* rng_state, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif  /* TINYNERF_B200_H_ */

// Anatomical-mask primitives for training batches: the scipy.ndimage / scipy.spatial / matplotlib.path pieces of the reference's
// modules/mask_generator.py on the GPU, bit-exact, for batches of 2-D slices:
//   label4        ndimage.label, default structure (4-connectivity), scipy's numbering      mask_generator.py:32,46,64,85
//   fill_holes    ndimage.binary_fill_holes, default structure                              mask_generator.py:69,90,242,310
//   detect_lung   thresholds + border margin + small-component removal                      mask_generator.py:11-52
//   detect_lung_vessels  component count / area test + (fill_holes - lung) & HU range       mask_generator.py:55-99
//   lung_hull / detect_mediastinum / detect_bone   ConvexHull vertices + Path.contains_points, region growing  mask_generator.py:100-311
// (the hull part is at the end of this file; matplotlib's rasterisation is restated, parity-unpinned).
//
// Connected components: union-find over the pixels of a slice (Komura / Playne-Hawick style).  Every foreground pixel
// starts as its own root; a pixel is united with its left and upper foreground neighbours by atomicMin on the parent
// links, which only ever decrease, so the root of a component ends up as its SMALLEST raster index whatever the thread
// interleaving -- the result is deterministic although the construction uses atomics.  scipy numbers components in the
// order their first pixel appears in a raster scan, i.e. by increasing root index: a per-slice exclusive scan over the
// "is a root" flags turns roots into scipy's labels.  All of it is integer work, bounded by HBM bandwidth (a few passes
// over 1-5 bytes per pixel); nothing here is GEMM-shaped.
#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kT = 256;

__device__ __forceinline__ int ld_link(const int* L, int i) { return reinterpret_cast<const volatile int*>(L)[i]; }

__device__ __forceinline__ int find_root(const int* L, int i) {
  int p = ld_link(L, i);
  while (p != i) {
    i = p;
    p = ld_link(L, i);
  }
  return i;
}

__device__ void unite(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = old == b;
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = old == a;
      a = old;
    } else {
      done = true;
    }
  }
}

// links: L[g] = g for foreground (mask != 0) pixels, -1 for background.  `invert` labels the background instead.
__global__ void __launch_bounds__(kT) ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, long long n, int invert) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const bool fg = (mask[g] != 0) != (invert != 0);
    L[g] = fg ? int(g) : -1;
  }
}

__global__ void __launch_bounds__(kT) ccl_merge_kernel(int* __restrict__ L, int H, int W, long long n) {
  pdl_prologue();
  const int HW = H * W;
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    if (ld_link(L, int(g)) < 0) continue;
    const int p = int(g % HW), y = p / W, x = p - y * W;
    if (x > 0 && ld_link(L, int(g) - 1) >= 0) unite(L, int(g), int(g) - 1);
    if (y > 0 && ld_link(L, int(g) - W) >= 0) unite(L, int(g), int(g) - W);
  }
}

// path compression to the root; optionally component sizes (integer atomics: order-independent) and "touches the slice
// border" flags per root
__global__ void __launch_bounds__(kT) ccl_flatten_kernel(int* __restrict__ L, int* __restrict__ size, uint8_t* __restrict__ border,
                                                         int H, int W, long long n) {
  pdl_prologue();
  const int HW = H * W;
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    if (ld_link(L, int(g)) < 0) continue;
    const int r = find_root(L, int(g));
    L[g] = r;                         // a link may be shortened while others still walk it: it stays inside the component
    if (size != nullptr) atomicAdd(&size[r], 1);
    if (border != nullptr) {
      const int p = int(g % HW), y = p / W, x = p - y * W;
      if (x == 0 || y == 0 || x == W - 1 || y == H - 1) border[r] = 1;
    }
  }
}

// ---- scipy numbering: rank of every root among the roots of its slice, in raster order ---------------------------------
// step 1: roots per chunk of kT pixels
__global__ void __launch_bounds__(kT) root_count_kernel(const int* __restrict__ L, int* __restrict__ chunk_count, int HW, int chunks) {
  pdl_prologue();
  const int b = blockIdx.y, c = blockIdx.x;
  const int p = c * kT + threadIdx.x;
  const long long g = (long long)b * HW + p;
  const int is_root = (p < HW && L[g] == int(g)) ? 1 : 0;
  const int n = __syncthreads_count(is_root);
  if (threadIdx.x == 0) chunk_count[b * chunks + c] = n;
}
// step 2: exclusive scan of the chunk counts of one slice (one CTA per slice), total = number of components
__global__ void __launch_bounds__(kT) chunk_scan_kernel(int* __restrict__ chunk_count, int* __restrict__ num, int chunks) {
  pdl_prologue();
  __shared__ int carry;
  __shared__ int warp_tot[kT / 32];
  int* cc = chunk_count + (long long)blockIdx.x * chunks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < chunks; base += kT) {
    const int i = base + threadIdx.x;
    const int v = i < chunks ? cc[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
    const int c0 = carry;
    if (i < chunks) cc[i] = c0 + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == kT - 1) carry = c0 + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && num != nullptr) num[blockIdx.x] = carry;
}
// step 3: rank[g] = 1 + (number of roots before g in its slice), for roots
__global__ void __launch_bounds__(kT) root_rank_kernel(const int* __restrict__ L, const int* __restrict__ chunk_off, int* __restrict__ rank,
                                                       int HW, int chunks) {
  pdl_prologue();
  __shared__ int warp_tot[kT / 32];
  const int b = blockIdx.y, c = blockIdx.x;
  const int p = c * kT + threadIdx.x;
  const long long g = (long long)b * HW + p;
  const int is_root = (p < HW && L[g] == int(g)) ? 1 : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, is_root);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  if (is_root) rank[g] = chunk_off[b * chunks + c] + woff + __popc(bal & ((1u << lane) - 1u)) + 1;
}
// step 4: labels[g] = rank[root(g)] (0 for background)
__global__ void __launch_bounds__(kT) relabel_kernel(const int* __restrict__ L, const int* __restrict__ rank, int* __restrict__ labels, long long n) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const int r = L[g];
    labels[g] = r < 0 ? 0 : rank[r];
  }
}

// ---- mask_generator.py pieces ---------------------------------------------------------------------------------------
// lung candidate: (lo <= hu <= hi) & (hu > -1000), border margin cleared   (mask_generator.py:13-29); body = hu > -1000
__global__ void __launch_bounds__(kT) lung_candidate_kernel(const float* __restrict__ hu, uint8_t* __restrict__ cand, uint8_t* __restrict__ body,
                                                            int H, int W, long long n, float lo, float hi, int margin) {
  pdl_prologue();
  const int HW = H * W;
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const float v = hu[g];
    const int p = int(g % HW), y = p / W, x = p - y * W;
    const bool bd = v > -1000.f;
    const bool inside = y >= margin && y < H - margin && x >= margin && x < W - margin;
    cand[g] = (bd && v >= lo && v <= hi && inside) ? 1 : 0;
    if (body != nullptr) body[g] = bd ? 1 : 0;
  }
}
// keep components of at least min_size pixels   (mask_generator.py:32-36)
__global__ void __launch_bounds__(kT) keep_large_kernel(const int* __restrict__ L, const int* __restrict__ size, uint8_t* __restrict__ out,
                                                        long long n, int min_size) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const int r = L[g];
    out[g] = (r >= 0 && size[r] >= min_size) ? 1 : 0;
  }
}
// filled = mask | (background component that does not touch the border)
__global__ void __launch_bounds__(kT) fill_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ L, const uint8_t* __restrict__ border,
                                                  uint8_t* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const int r = L[g];     // links of the BACKGROUND labelling: r < 0 on foreground pixels
    out[g] = (mask[g] != 0 || (r >= 0 && border[r] == 0)) ? 1 : 0;
  }
}
// per-slice pixel counts of two uint8 masks (integer atomics)
__global__ void __launch_bounds__(kT) area_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int* __restrict__ area_a,
                                                  int* __restrict__ area_b, int HW) {
  pdl_prologue();
  const int s = blockIdx.y;
  int ca = 0, cb = 0;
  for (int p = blockIdx.x * kT + threadIdx.x; p < HW; p += gridDim.x * kT) {
    ca += a[(long long)s * HW + p] != 0;
    cb += b[(long long)s * HW + p] != 0;
  }
  // warp reduce, then one integer atomic per warp (order-independent)
  for (int o = 16; o > 0; o >>= 1) {
    ca += __shfl_xor_sync(0xffffffffu, ca, o);
    cb += __shfl_xor_sync(0xffffffffu, cb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ca) atomicAdd(&area_a[s], ca);
    if (cb) atomicAdd(&area_b[s], cb);
  }
}
// vessel = (filled - lung) & (lo <= hu <= hi) where the slice passes the reference's plausibility test
// (>= 2 lung components, body_area > 0, lung_area / body_area >= 0.1 in float64)   (mask_generator.py:63-76)
__global__ void __launch_bounds__(kT) vessel_kernel(const float* __restrict__ hu, const uint8_t* __restrict__ lung, const uint8_t* __restrict__ filled,
                                                    const int* __restrict__ num_regions, const int* __restrict__ body_area,
                                                    const int* __restrict__ lung_area, uint8_t* __restrict__ out, int HW, long long n, float lo,
                                                    float hi) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const int s = int(g / HW);
    const bool ok = num_regions[s] >= 2 && body_area[s] > 0 && (double(lung_area[s]) / double(body_area[s])) >= 0.1;
    const float v = hu[g];
    out[g] = (ok && filled[g] != 0 && lung[g] == 0 && v >= lo && v <= hi) ? 1 : 0;
  }
}

// ---- second half: convex hull of the lung pixels of a slice (mask_generator.py:115-127,204-216) ------------------------------
// scipy.spatial.ConvexHull(np.argwhere(lung == 1)).vertices are the strictly convex corners of the pixel set in
// counter-clockwise order (coordinates (row, col)).  Only the leftmost and rightmost lung pixel of every row can be corners, so:
// per-row extreme columns (shared-memory atomics), then Andrew's monotone chain over those <= 2H candidates by one thread
// (already sorted by row, then column; collinear points are popped, as qhull reports only true vertices), counter-clockwise,
// integer cross products.  nverts = 0 marks the reference's fallback branches (fewer than 3 pixels / QhullError on a
// degenerate set): the callers then use the lung mask itself as "hull".
__global__ void __launch_bounds__(kT) hull_build_kernel(const uint8_t* __restrict__ lung, int* __restrict__ cand, int* __restrict__ verts,
                                                        int* __restrict__ nverts, int H, int W, int maxV) {
  pdl_prologue();
  extern __shared__ int hull_sm[];      // [H] min column, [H] max column
  int* minc = hull_sm;
  int* maxc = hull_sm + H;
  const int s = blockIdx.x;
  for (int r = threadIdx.x; r < H; r += kT) { minc[r] = 0x7fffffff; maxc[r] = -1; }
  __syncthreads();
  const uint8_t* m = lung + (long long)s * H * W;
  for (int p = threadIdx.x; p < H * W; p += kT)
    if (m[p] != 0) {
      const int y = p / W, x = p - y * W;
      atomicMin(&minc[y], x);
      atomicMax(&maxc[y], x);
    }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int* c = cand + (long long)s * maxV * 2;
  int* v = verts + (long long)s * maxV * 2;
  int nc = 0;
  for (int r = 0; r < H; ++r)
    if (maxc[r] >= 0) {
      c[2 * nc] = r; c[2 * nc + 1] = minc[r]; ++nc;
      if (maxc[r] != minc[r]) { c[2 * nc] = r; c[2 * nc + 1] = maxc[r]; ++nc; }
    }
  auto cross = [&](int o, int a, const int* b) {   // (v[a] - v[o]) x (b - v[o])
    return (long long)(v[2 * a] - v[2 * o]) * (b[1] - v[2 * o + 1]) - (long long)(v[2 * a + 1] - v[2 * o + 1]) * (b[0] - v[2 * o]);
  };
  int k = 0;
  for (int i = 0; i < nc; ++i) {                   // lower chain
    while (k >= 2 && cross(k - 2, k - 1, c + 2 * i) <= 0) --k;
    v[2 * k] = c[2 * i]; v[2 * k + 1] = c[2 * i + 1]; ++k;
  }
  const int t = k + 1;
  for (int i = nc - 2; i >= 0; --i) {              // upper chain
    while (k >= t && cross(k - 2, k - 1, c + 2 * i) <= 0) --k;
    v[2 * k] = c[2 * i]; v[2 * k + 1] = c[2 * i + 1]; ++k;
  }
  const int nv = k - 1;                            // the last point repeats the first
  nverts[s] = (nc >= 3 && nv >= 3) ? nv : 0;
}

// hull_mask = matplotlib.path.Path(vertices).contains_points(every pixel) -- the crossings test of matplotlib's
// point_in_path_impl with its inequality conventions (points exactly on the boundary depend on them; PARITY UNPINNED,
// matplotlib is absent here; restated in oracle.path_contains_points).  Path coordinates: x = row, y = column.  Integer
// arithmetic (all products < 2^31).  nverts == 0: hull_mask = lung (the reference's fallback).
__global__ void __launch_bounds__(kT) hull_raster_kernel(const int* __restrict__ verts, const int* __restrict__ nverts,
                                                         const uint8_t* __restrict__ lung, uint8_t* __restrict__ hull, int H, int W, int maxV) {
  pdl_prologue();
  extern __shared__ int hull_sm[];      // [nv][2]
  const int s = blockIdx.y, nv = nverts[s];
  const int* v = verts + (long long)s * maxV * 2;
  for (int i = threadIdx.x; i < 2 * nv; i += kT) hull_sm[i] = v[i];
  __syncthreads();
  for (int p = blockIdx.x * kT + threadIdx.x; p < H * W; p += gridDim.x * kT) {
    const long long g = (long long)s * H * W + p;
    if (nv == 0) { hull[g] = lung[g]; continue; }
    const int tx = p / W, ty = p - tx * W;
    int inside = 0;
    int x0 = hull_sm[2 * (nv - 1)], y0 = hull_sm[2 * (nv - 1) + 1];
    for (int i = 0; i < nv; ++i) {
      const int x1 = hull_sm[2 * i], y1 = hull_sm[2 * i + 1];
      const bool f0 = y0 >= ty, f1 = y1 >= ty;
      if (f0 != f1 && (((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1)) inside ^= 1;
      x0 = x1; y0 = y1;
    }
    hull[g] = uint8_t(inside);
  }
}

// bone candidates (hu >= threshold) & (hu > -1000)   (mask_generator.py:177-179)
__global__ void __launch_bounds__(kT) bone_candidate_kernel(const float* __restrict__ hu, uint8_t* __restrict__ cand, long long n, float thr) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const float v = hu[g];
    cand[g] = (v >= thr && v > -1000.f) ? 1 : 0;
  }
}
// reduced = cand & ~(hull & ~lung & rows above the preserved spine rows) on slices that pass the plausibility test and have a
// real hull   (mask_generator.py:200-228)
__global__ void __launch_bounds__(kT) bone_reduce_kernel(const uint8_t* __restrict__ cand, const uint8_t* __restrict__ hull,
                                                         const uint8_t* __restrict__ lung, const int* __restrict__ num_regions,
                                                         const int* __restrict__ body_area, const int* __restrict__ lung_area,
                                                         const int* __restrict__ nverts, uint8_t* __restrict__ reduced, int H, int W,
                                                         long long n, int spine_start) {
  pdl_prologue();
  const int HW = H * W;
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT) {
    const int s = int(g / HW), y = int(g % HW) / W;
    const bool ok = num_regions[s] >= 2 && body_area[s] > 0 && (double(lung_area[s]) / double(body_area[s])) >= 0.1 && nverts[s] > 0;
    const bool region = ok && hull[g] != 0 && lung[g] == 0 && y < spine_start;
    reduced[g] = (cand[g] != 0 && !region) ? 1 : 0;
  }
}
// region growing (mask_generator.py:230-246): a component of the candidates is kept whole when any of its pixels survived
__global__ void __launch_bounds__(kT) touch_kernel(const int* __restrict__ L, const uint8_t* __restrict__ reduced, int* __restrict__ flag, long long n) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT)
    if (reduced[g] != 0 && L[g] >= 0) flag[L[g]] = 1;     // every writer stores the same value
}
__global__ void __launch_bounds__(kT) bone_select_kernel(const uint8_t* __restrict__ cand, const int* __restrict__ L, const int* __restrict__ flag,
                                                         uint8_t* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long g = (long long)blockIdx.x * kT + threadIdx.x; g < n; g += (long long)gridDim.x * kT)
    out[g] = (cand[g] != 0 && L[g] >= 0 && flag[L[g]] != 0) ? 1 : 0;
}

int ew_grid(long long n) {
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
  const long long b = (n + kT - 1) / kT;
  return int(b < cap ? (b > 0 ? b : 1) : cap);
}

struct Scratch {
  int *L, *aux, *chunk, *num, *area_a, *area_b, *hull_cand, *hull_verts, *hull_n;
  uint8_t *u8a, *u8b, *u8c;
  size_t total;
};
inline int hull_max_verts(int H) { return 2 * H + 4; }
Scratch carve(void* base, int B, int H, int W) {
  const size_t n = size_t(B) * H * W;
  const size_t chunks = (size_t(H) * W + kT - 1) / kT;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* q = p ? p + off : nullptr;
    off = (off + bytes + 255) / 256 * 256;
    return q;
  };
  Scratch s;
  s.L = reinterpret_cast<int*>(take(n * 4));
  s.aux = reinterpret_cast<int*>(take(n * 4));          // component sizes / ranks
  s.chunk = reinterpret_cast<int*>(take(size_t(B) * chunks * 4));
  s.num = reinterpret_cast<int*>(take(size_t(B) * 4));
  s.area_a = reinterpret_cast<int*>(take(size_t(B) * 4));
  s.area_b = reinterpret_cast<int*>(take(size_t(B) * 4));
  s.u8a = take(n);
  s.u8b = take(n);
  s.u8c = take(n);
  s.hull_cand = reinterpret_cast<int*>(take(size_t(B) * hull_max_verts(H) * 2 * 4));
  s.hull_verts = reinterpret_cast<int*>(take(size_t(B) * hull_max_verts(H) * 2 * 4));
  s.hull_n = reinterpret_cast<int*>(take(size_t(B) * 4));
  s.total = off;
  return s;
}

int check_shape(const char* who, int B, int H, int W) {
  DUCOSY_CHECK(B > 0 && H > 0 && W > 0, DUCOSY_ERR_SHAPE, "%s: bad shape", who);
  DUCOSY_CHECK((long long)B * H * W < (1LL << 31) && B <= 65535, DUCOSY_ERR_SHAPE, "%s: batch too large for 32-bit pixel indices", who);
  return 0;
}

// union-find labelling of `mask` (or of its complement) into links; optional sizes / border flags (zeroed here)
int run_ccl(const uint8_t* mask, int invert, int* L, int* size, uint8_t* border, int B, int H, int W, cudaStream_t st) {
  const long long n = (long long)B * H * W;
  if (size != nullptr) cudaMemsetAsync(size, 0, size_t(n) * 4, st);
  if (border != nullptr) cudaMemsetAsync(border, 0, size_t(n), st);
  pdl(ccl_init_kernel, ew_grid(n), kT, 0, st)(mask, L, n, invert);
  pdl(ccl_merge_kernel, ew_grid(n), kT, 0, st)(L, H, W, n);
  pdl(ccl_flatten_kernel, ew_grid(n), kT, 0, st)(L, size, border, H, W, n);
  return check_launch("ccl kernels");
}

// number of components per slice (and optionally scipy labels) from flattened links
int run_count(const int* L, int* chunk, int* num, int* rank, int* labels, int B, int H, int W, cudaStream_t st) {
  const int HW = H * W, chunks = (HW + kT - 1) / kT;
  pdl(root_count_kernel, dim3(chunks, B), kT, 0, st)(L, chunk, HW, chunks);
  pdl(chunk_scan_kernel, B, kT, 0, st)(chunk, num, chunks);
  if (labels != nullptr) {
    const long long n = (long long)B * HW;
    pdl(root_rank_kernel, dim3(chunks, B), kT, 0, st)(L, chunk, rank, HW, chunks);
    pdl(relabel_kernel, ew_grid(n), kT, 0, st)(L, rank, labels, n);
  }
  return check_launch("component count kernels");
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" size_t ducosy_masks_scratch_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return carve(nullptr, B, H, W).total;
}

extern "C" int ducosy_label4(const uint8_t* mask, int32_t* labels, int32_t* num_features, int B, int H, int W, void* scratch,
                             size_t scratch_bytes, ducosy_stream_t stream) {
  DUCOSY_CHECK(mask && labels && scratch, DUCOSY_ERR_ARG, "label4: null pointer");
  DUCOSY_TRY(check_shape("label4", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "label4: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUCOSY_TRY(run_ccl(mask, 0, s.L, nullptr, nullptr, B, H, W, st));
  DUCOSY_TRY(run_count(s.L, s.chunk, s.num, s.aux, labels, B, H, W, st));
  if (num_features != nullptr) cudaMemcpyAsync(num_features, s.num, size_t(B) * 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("label4");
}

extern "C" int ducosy_binary_fill_holes(const uint8_t* mask, uint8_t* out, int B, int H, int W, void* scratch, size_t scratch_bytes,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(mask && out && scratch, DUCOSY_ERR_ARG, "binary_fill_holes: null pointer");
  DUCOSY_TRY(check_shape("binary_fill_holes", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "binary_fill_holes: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  DUCOSY_TRY(run_ccl(mask, 1, s.L, nullptr, s.u8a, B, H, W, st));
  pdl(fill_kernel, ew_grid(n), kT, 0, st)(mask, s.L, s.u8a, out, n);
  return check_launch("fill_kernel");
}

extern "C" int ducosy_detect_lung(const float* hu, uint8_t* lung_mask, int B, int H, int W, float lung_lower, float lung_upper,
                                  int min_size, int border_margin, void* scratch, size_t scratch_bytes, ducosy_stream_t stream) {
  DUCOSY_CHECK(hu && lung_mask && scratch, DUCOSY_ERR_ARG, "detect_lung: null pointer");
  DUCOSY_TRY(check_shape("detect_lung", B, H, W));
  DUCOSY_CHECK(border_margin >= 0 && min_size >= 0, DUCOSY_ERR_ARG, "detect_lung: negative margin / size");
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "detect_lung: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  pdl(lung_candidate_kernel, ew_grid(n), kT, 0, st)(hu, s.u8a, nullptr, H, W, n, lung_lower, lung_upper, border_margin);
  DUCOSY_TRY(check_launch("lung_candidate_kernel"));
  DUCOSY_TRY(run_ccl(s.u8a, 0, s.L, s.aux, nullptr, B, H, W, st));
  pdl(keep_large_kernel, ew_grid(n), kT, 0, st)(s.L, s.aux, lung_mask, n, min_size);
  return check_launch("keep_large_kernel");
}

extern "C" int ducosy_detect_lung_vessels(const float* hu, const uint8_t* lung_mask, uint8_t* vessel_mask, int B, int H, int W,
                                          float vessel_lower, float vessel_upper, void* scratch, size_t scratch_bytes,
                                          ducosy_stream_t stream) {
  DUCOSY_CHECK(hu && lung_mask && vessel_mask && scratch, DUCOSY_ERR_ARG, "detect_lung_vessels: null pointer");
  DUCOSY_TRY(check_shape("detect_lung_vessels", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "detect_lung_vessels: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  const int HW = H * W;
  // number of lung components per slice
  DUCOSY_TRY(run_ccl(lung_mask, 0, s.L, nullptr, nullptr, B, H, W, st));
  DUCOSY_TRY(run_count(s.L, s.chunk, s.num, nullptr, nullptr, B, H, W, st));
  // body mask and the two areas
  pdl(lung_candidate_kernel, ew_grid(n), kT, 0, st)(hu, s.u8c, s.u8b, H, W, n, 0.f, -1.f, 0);   // only the body mask (u8b) is used
  cudaMemsetAsync(s.area_a, 0, size_t(B) * 4, st);
  cudaMemsetAsync(s.area_b, 0, size_t(B) * 4, st);
  pdl(area_kernel, dim3(std::min((HW + kT - 1) / kT, 64), B), kT, 0, st)(s.u8b, lung_mask, s.area_a, s.area_b, HW);
  DUCOSY_TRY(check_launch("area_kernel"));
  // filled lung
  DUCOSY_TRY(run_ccl(lung_mask, 1, s.L, nullptr, s.u8a, B, H, W, st));
  pdl(fill_kernel, ew_grid(n), kT, 0, st)(lung_mask, s.L, s.u8a, s.u8c, n);
  pdl(vessel_kernel, ew_grid(n), kT, 0, st)(hu, lung_mask, s.u8c, s.num, s.area_a, s.area_b, vessel_mask, HW, n, vessel_lower, vessel_upper);
  return check_launch("vessel_kernel");
}

namespace {
// plausibility inputs of a slice (component count of the lung mask, body / lung areas) and the rasterised hull into s.u8c;
// s.u8b receives the body mask.  Uses s.L, s.chunk, s.num, s.area_*, s.hull_*.
int run_hull(const float* hu, const uint8_t* lung_mask, const ducosy::Scratch& s, int B, int H, int W, cudaStream_t st) {
  const long long n = (long long)B * H * W;
  const int HW = H * W;
  DUCOSY_CHECK(size_t(2) * H * sizeof(int) <= 48 * 1024 && size_t(2) * hull_max_verts(H) * sizeof(int) <= 48 * 1024, DUCOSY_ERR_SHAPE,
               "lung hull: H = %d too large", H);
  DUCOSY_TRY(run_ccl(lung_mask, 0, s.L, nullptr, nullptr, B, H, W, st));
  DUCOSY_TRY(run_count(s.L, s.chunk, s.num, nullptr, nullptr, B, H, W, st));
  pdl(lung_candidate_kernel, ew_grid(n), kT, 0, st)(hu, s.u8c, s.u8b, H, W, n, 0.f, -1.f, 0);   // only the body mask (u8b) is used
  cudaMemsetAsync(s.area_a, 0, size_t(B) * 4, st);
  cudaMemsetAsync(s.area_b, 0, size_t(B) * 4, st);
  pdl(area_kernel, dim3(std::min((HW + kT - 1) / kT, 64), B), kT, 0, st)(s.u8b, lung_mask, s.area_a, s.area_b, HW);
  DUCOSY_TRY(check_launch("area_kernel"));
  const int maxV = hull_max_verts(H);
  pdl(hull_build_kernel, B, kT, size_t(2) * H * sizeof(int), st)(lung_mask, s.hull_cand, s.hull_verts, s.hull_n, H, W, maxV);
  DUCOSY_TRY(check_launch("hull_build_kernel"));
  pdl(hull_raster_kernel, dim3(std::min((HW + kT - 1) / kT, 128), B), kT, size_t(2) * maxV * sizeof(int), st)(
      s.hull_verts, s.hull_n, lung_mask, s.u8c, H, W, maxV);
  return check_launch("hull_raster_kernel");
}
}  // namespace

extern "C" int ducosy_lung_hull(const uint8_t* lung_mask, uint8_t* hull_mask, int32_t* verts, int32_t* nverts, int B, int H, int W,
                                void* scratch, size_t scratch_bytes, ducosy_stream_t stream) {
  DUCOSY_CHECK(lung_mask && hull_mask && scratch, DUCOSY_ERR_ARG, "lung_hull: null pointer");
  DUCOSY_TRY(check_shape("lung_hull", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "lung_hull: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  DUCOSY_CHECK(size_t(2) * hull_max_verts(H) * sizeof(int) <= 48 * 1024, DUCOSY_ERR_SHAPE, "lung_hull: H = %d too large", H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int maxV = hull_max_verts(H), HW = H * W;
  pdl(hull_build_kernel, B, kT, size_t(2) * H * sizeof(int), st)(lung_mask, s.hull_cand, s.hull_verts, s.hull_n, H, W, maxV);
  DUCOSY_TRY(check_launch("hull_build_kernel"));
  pdl(hull_raster_kernel, dim3(std::min((HW + kT - 1) / kT, 128), B), kT, size_t(2) * maxV * sizeof(int), st)(
      s.hull_verts, s.hull_n, lung_mask, hull_mask, H, W, maxV);
  DUCOSY_TRY(check_launch("hull_raster_kernel"));
  if (verts != nullptr) cudaMemcpyAsync(verts, s.hull_verts, size_t(B) * maxV * 2 * 4, cudaMemcpyDeviceToDevice, st);
  if (nverts != nullptr) cudaMemcpyAsync(nverts, s.hull_n, size_t(B) * 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("lung_hull");
}

extern "C" int ducosy_detect_mediastinum(const float* hu, const uint8_t* lung_mask, uint8_t* mediastinum_mask, int B, int H, int W,
                                         float lower, float upper, void* scratch, size_t scratch_bytes, ducosy_stream_t stream) {
  DUCOSY_CHECK(hu && lung_mask && mediastinum_mask && scratch, DUCOSY_ERR_ARG, "detect_mediastinum: null pointer");
  DUCOSY_TRY(check_shape("detect_mediastinum", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "detect_mediastinum: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  DUCOSY_TRY(run_hull(hu, lung_mask, s, B, H, W, st));
  // (hull - lung) & HU range on slices that pass the plausibility test: the vessel kernel with the hull in place of the filled lung
  pdl(vessel_kernel, ew_grid(n), kT, 0, st)(hu, lung_mask, s.u8c, s.num, s.area_a, s.area_b, mediastinum_mask, H * W, n, lower, upper);
  return check_launch("vessel_kernel(mediastinum)");
}

extern "C" int ducosy_detect_bone(const float* hu, const uint8_t* lung_mask, uint8_t* bone_mask, int B, int H, int W, float bone_threshold,
                                  int spine_start_row, void* scratch, size_t scratch_bytes, ducosy_stream_t stream) {
  DUCOSY_CHECK(hu && lung_mask && bone_mask && scratch, DUCOSY_ERR_ARG, "detect_bone: null pointer");
  DUCOSY_TRY(check_shape("detect_bone", B, H, W));
  const Scratch s = carve(scratch, B, H, W);
  DUCOSY_CHECK(scratch_bytes >= s.total, DUCOSY_ERR_WORKSPACE, "detect_bone: scratch %zu < required %zu bytes", scratch_bytes, s.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  DUCOSY_TRY(run_hull(hu, lung_mask, s, B, H, W, st));                      // hull in u8c, counts / areas / nverts filled
  pdl(bone_candidate_kernel, ew_grid(n), kT, 0, st)(hu, s.u8a, n, bone_threshold);
  pdl(bone_reduce_kernel, ew_grid(n), kT, 0, st)(s.u8a, s.u8c, lung_mask, s.num, s.area_a, s.area_b, s.hull_n, s.u8b, H, W, n, spine_start_row);
  DUCOSY_TRY(check_launch("bone_reduce_kernel"));
  DUCOSY_TRY(run_ccl(s.u8a, 0, s.L, nullptr, nullptr, B, H, W, st));        // components of the candidates
  cudaMemsetAsync(s.aux, 0, size_t(n) * 4, st);
  pdl(touch_kernel, ew_grid(n), kT, 0, st)(s.L, s.u8b, s.aux, n);
  pdl(bone_select_kernel, ew_grid(n), kT, 0, st)(s.u8a, s.L, s.aux, s.u8c, n);
  DUCOSY_TRY(check_launch("bone_select_kernel"));
  DUCOSY_TRY(run_ccl(s.u8c, 1, s.L, nullptr, s.u8a, B, H, W, st));          // fill holes
  pdl(fill_kernel, ew_grid(n), kT, 0, st)(s.u8c, s.L, s.u8a, bone_mask, n);
  return check_launch("fill_kernel(bone)");
}

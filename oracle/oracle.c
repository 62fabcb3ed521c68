/*
 * oracle.c — CPU restatement of the reference's aggregation path.  TEST
 * INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the checker and the
 * reported CPU baseline.  Nothing in gnn-ops-benchmark_b200/ (the product)
 * may link, import or call it.
 *
 * PARITY STATUS.  The arithmetic the reference times lives in pinned,
 * un-vendored dependencies that are absent from /root/reference and not
 * installable here: torch-scatter==2.0.9, torch-sparse==0.6.12,
 * torch==1.11.0 (requirements.txt:209-213).  The reference holds no tests,
 * golden vectors or fixtures for them, so for the
 * torch_scatter / torch_sparse entry points this oracle is "PARITY UNPINNED":
 * it restates the published upstream algorithm and is cross-checked against
 * independent formulations (torch scatter_reduce_/scatter_add_/index_add_,
 * scipy.sparse) in tests/test_oracle.py.  For the NATIVE torch ops the
 * reference calls directly (index_add_, index_select, scatter_(multiply),
 * sparse.mm, sort, Tensor.coalesce) it is pinned against golden vectors made
 * by executing the reference's own op functions (tests/golden/make_golden.py).
 *
 * Every function names the reference call site it follows (file:line relative
 * to the reference root) and the upstream routine it restates.
 *
 * Values are float32 on this side; half/bfloat16 inputs are widened by the
 * Python wrapper and results rounded once (the "fp32-accumulate, round once"
 * contract of DESIGN.md).  Indices are int64 as in the reference.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { RED_SUM = 0, RED_MEAN = 1, RED_MUL = 2, RED_MIN = 3, RED_MAX = 4 };

/*
 * torch_scatter.scatter_{sum,mul,min,max,mean} — call sites
 * op_bm_scripts/benchmark_scatter_add.py:18, _mean.py:17, _max.py:17, _min.py:17,
 * benchmark_scatter_multiply.py:44 (native spelling of mul).
 * Restates torch-scatter 2.0.9 csrc/cpu/scatter_cpu.cpp: src viewed as
 * [B, E, K]; sequential loops b, e, k; out pre-filled with the reduction's
 * identity (0 / 1 / lowest / max); MIN/MAX update only on a STRICT compare, so
 * the lowest e among equal values wins and NaNs never win; arg pre-filled
 * with E; afterwards entries still equal to the identity are set to 0
 * (out.masked_fill_).  MEAN = SUM, then divide by the count clamped to >= 1
 * (python scatter_mean).  index is either full-shape [B, E, K]
 * (index_is_1d = 0) or a 1-D vector of length E broadcast over b and k.
 * lowest/highest are passed in so half types use their own finite range.
 * Entries with index outside [0, N) are skipped.
 */
void oracle_scatter(const float* src, const int64_t* index, int index_is_1d, int64_t B, int64_t E,
                    int64_t K, int64_t N, int reduce, float lowest, float highest, float* out,
                    int64_t* arg /* may be NULL */) {
  const int64_t n_out = B * N * K;
  float init = 0.f;
  if (reduce == RED_MUL) init = 1.f;
  if (reduce == RED_MAX) init = lowest;
  if (reduce == RED_MIN) init = highest;
  float* cnt = NULL;
  for (int64_t i = 0; i < n_out; ++i) out[i] = init;
  if (arg)
    for (int64_t i = 0; i < n_out; ++i) arg[i] = E;
  if (reduce == RED_MEAN) cnt = (float*)calloc((size_t)(n_out > 0 ? n_out : 1), sizeof(float));
  for (int64_t b = 0; b < B; ++b)
    for (int64_t e = 0; e < E; ++e)
      for (int64_t k = 0; k < K; ++k) {
        const int64_t i = (b * E + e) * K + k;
        const int64_t idx = index_is_1d ? index[e] : index[i];
        if (idx < 0 || idx >= N) continue;
        const int64_t t = (b * N + idx) * K + k;
        const float v = src[i];
        switch (reduce) {
          case RED_SUM: out[t] += v; break;
          case RED_MEAN: out[t] += v; cnt[t] += 1.f; break;
          case RED_MUL: out[t] *= v; break;
          case RED_MIN: if (v < out[t]) { out[t] = v; if (arg) arg[t] = e; } break;
          case RED_MAX: if (v > out[t]) { out[t] = v; if (arg) arg[t] = e; } break;
        }
      }
  if (reduce == RED_MEAN) {
    for (int64_t i = 0; i < n_out; ++i) out[i] = out[i] / (cnt[i] < 1.f ? 1.f : cnt[i]);
    free(cnt);
  }
  if (reduce == RED_MIN || reduce == RED_MAX)
    for (int64_t i = 0; i < n_out; ++i)
      if (out[i] == init) out[i] = 0.f;
}

/*
 * Fused message passing: scatter(x.index_select(0, src_ids), dst_ids, dim=0,
 * dim_size=N, reduce) — PyG MessagePassing.propagate as reached from
 * graph_benchmark/models/ptg_models.py:238-258, and the gather half timed by
 * op_bm_scripts/benchmark_fused_index_select_reduce.py:12-15.  Restated as
 * the un-fused sequence the reference executes: materialise the messages in
 * edge order, then oracle_scatter's loop (here without the temporary).
 */
void oracle_gather_scatter(const float* x, int64_t F, const int64_t* src_ids,
                           const int64_t* dst_ids, int64_t E, int64_t N, int reduce, float lowest,
                           float highest, float* out, int64_t* arg) {
  const int64_t n_out = N * F;
  float init = 0.f;
  if (reduce == RED_MUL) init = 1.f;
  if (reduce == RED_MAX) init = lowest;
  if (reduce == RED_MIN) init = highest;
  for (int64_t i = 0; i < n_out; ++i) out[i] = init;
  if (arg)
    for (int64_t i = 0; i < n_out; ++i) arg[i] = E;
  int64_t* cnt = (int64_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(int64_t));
  for (int64_t e = 0; e < E; ++e) {
    const int64_t d = dst_ids[e];
    if (d < 0 || d >= N) continue;
    const float* xr = x + src_ids[e] * F;
    float* o = out + d * F;
    cnt[d] += 1;
    for (int64_t k = 0; k < F; ++k) {
      const float v = xr[k];
      switch (reduce) {
        case RED_SUM: case RED_MEAN: o[k] += v; break;
        case RED_MUL: o[k] *= v; break;
        case RED_MIN: if (v < o[k]) { o[k] = v; if (arg) arg[d * F + k] = e; } break;
        case RED_MAX: if (v > o[k]) { o[k] = v; if (arg) arg[d * F + k] = e; } break;
      }
    }
  }
  if (reduce == RED_MEAN)
    for (int64_t d = 0; d < N; ++d) {
      const float c = (float)(cnt[d] < 1 ? 1 : cnt[d]);
      for (int64_t k = 0; k < F; ++k) out[d * F + k] = out[d * F + k] / c;
    }
  if (reduce == RED_MIN || reduce == RED_MAX)
    for (int64_t i = 0; i < n_out; ++i)
      if (out[i] == init) out[i] = 0.f;
  free(cnt);
}

/*
 * input.index_add_(dim, index, source) on [B, E, K]-viewed tensors —
 * op_bm_scripts/benchmark_native_index_add_.py:13-16 and
 * benchmark_fused_index_add_reduce.py:12-15.  ATen index_add semantics:
 * inout[b, index[e], k] += source[b, e, k], sequential in e.
 */
void oracle_index_add(float* inout, const float* source, const int64_t* index, int64_t B, int64_t E,
                      int64_t K, int64_t N) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t e = 0; e < E; ++e) {
      const int64_t idx = index[e];
      if (idx < 0 || idx >= N) continue;
      for (int64_t k = 0; k < K; ++k) inout[(b * N + idx) * K + k] += source[(b * E + e) * K + k];
    }
}

/*
 * CSR SpMM with reductions — torch.sparse.mm at
 * op_bm_scripts/benchmark_sparse_spmm.py:12-14 and torch-sparse 0.6.12
 * csrc/cpu/spmm_cpu.cpp (spmm_{sum,mean,min,max}): per row, sequential over
 * its non-zeros; MIN/MAX arg = nnz position (sentinel nnz), empty rows → 0;
 * MEAN divides by the row length clamped to 1.  value may be NULL (all ones).
 */
void oracle_spmm_csr(const int64_t* rowptr, const int64_t* col, const float* value, const float* mat,
                     int64_t M, int64_t F, int reduce, float lowest, float highest, float* out,
                     int64_t* arg) {
  const int64_t nnz = rowptr[M];
  for (int64_t r = 0; r < M; ++r) {
    const int64_t kb = rowptr[r], ke = rowptr[r + 1];
    for (int64_t f = 0; f < F; ++f) {
      float a = 0.f;
      int64_t ae = nnz;
      if (reduce == RED_MAX) a = lowest;
      if (reduce == RED_MIN) a = highest;
      for (int64_t k = kb; k < ke; ++k) {
        const float v = (value ? value[k] : 1.f) * mat[col[k] * F + f];
        if (reduce == RED_SUM || reduce == RED_MEAN) a += v;
        else if (reduce == RED_MIN) { if (v < a) { a = v; ae = k; } }
        else if (reduce == RED_MAX) { if (v > a) { a = v; ae = k; } }
      }
      if (reduce == RED_MEAN) a = a / (float)(ke - kb < 1 ? 1 : ke - kb);
      if ((reduce == RED_MIN || reduce == RED_MAX) && ae == nnz) a = 0.f;
      out[r * F + f] = a;
      if (arg) arg[r * F + f] = ae;
    }
  }
}

/* ------------------------------------------------------------------ sorting */
typedef struct { uint64_t key; int64_t pos; } kv_t;
static void merge_sort_kv(kv_t* a, kv_t* tmp, int64_t n) {
  for (int64_t w = 1; w < n; w *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * w) {
      int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int64_t i = lo, j = mid, o = lo;
      while (i < mid && j < hi) tmp[o++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
      while (i < mid) tmp[o++] = a[i++];
      while (j < hi) tmp[o++] = a[j++];
    }
    memcpy(a, tmp, (size_t)n * sizeof(kv_t));
  }
}

/*
 * torch_sparse.coalesce(index, value, m, n, op) — call site
 * op_bm_scripts/benchmark_sparse_coalesce.py:35-37.  Restates torch-sparse
 * 0.6.12 storage.py/coalesce.py: key = row*n + col; sort by key (a STABLE
 * merge sort here, so duplicates keep input order — upstream's argsort leaves
 * that order unspecified, hence value tolerance rather than bit-exactness);
 * mask = key[k] > key[k-1]; unique (row, col) kept; values of a run reduced
 * with op (SUM/MEAN/MIN/MAX/MUL) sequentially.  Returns the merged count.
 * Passing (col, row, n, m) restates torch_sparse.transpose.
 */
int64_t oracle_coalesce(const int64_t* row, const int64_t* col, const float* value, int64_t K,
                        int64_t E, int64_t m, int64_t n, int reduce, int64_t* out_row,
                        int64_t* out_col, float* out_value) {
  (void)m;
  if (E == 0) return 0;
  kv_t* a = (kv_t*)malloc((size_t)E * sizeof(kv_t));
  kv_t* tmp = (kv_t*)malloc((size_t)E * sizeof(kv_t));
  for (int64_t e = 0; e < E; ++e) {
    a[e].key = (uint64_t)row[e] * (uint64_t)n + (uint64_t)col[e];
    a[e].pos = e;
  }
  merge_sort_kv(a, tmp, E);
  int64_t u = -1, run = 0;
  for (int64_t k = 0; k < E; ++k) {
    const int head = (k == 0) || (a[k].key != a[k - 1].key);
    if (head) {
      if (u >= 0 && reduce == RED_MEAN && value)
        for (int64_t j = 0; j < K; ++j) out_value[u * K + j] /= (float)run;
      ++u;
      run = 0;
      out_row[u] = (int64_t)(a[k].key / (uint64_t)n);
      out_col[u] = (int64_t)(a[k].key % (uint64_t)n);
    }
    ++run;
    if (value)
      for (int64_t j = 0; j < K; ++j) {
        const float v = value[a[k].pos * K + j];
        float* o = &out_value[u * K + j];
        if (head) *o = v;
        else if (reduce == RED_SUM || reduce == RED_MEAN) *o += v;
        else if (reduce == RED_MUL) *o *= v;
        else if (reduce == RED_MIN) *o = v < *o ? v : *o;
        else *o = v > *o ? v : *o;
      }
  }
  if (reduce == RED_MEAN && value)
    for (int64_t j = 0; j < K; ++j) out_value[u * K + j] /= (float)run;
  free(a);
  free(tmp);
  return u + 1;
}

/*
 * torch.sort(input, dim, stable=True) for float32 viewed as
 * [outer, len, inner] — op_bm_scripts/benchmark_native_sort.py:28-30.
 * ATen comparison semantics: a < b with every NaN greater than any number
 * (NaNs last, in input order) and -0.0 == +0.0 (input order kept); descending
 * uses a > b with the same stability.
 */
static uint32_t float_order_key(float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  if (f != f) return 0xffffffffu;
  if (b == 0x80000000u) b = 0;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
void oracle_sort_f32(const float* in, int64_t outer, int64_t len, int64_t inner, int descending,
                     float* out_values, int64_t* out_index) {
  kv_t* a = (kv_t*)malloc((size_t)(len > 0 ? len : 1) * sizeof(kv_t));
  kv_t* tmp = (kv_t*)malloc((size_t)(len > 0 ? len : 1) * sizeof(kv_t));
  for (int64_t o = 0; o < outer; ++o)
    for (int64_t i = 0; i < inner; ++i) {
      for (int64_t j = 0; j < len; ++j) {
        uint32_t k = float_order_key(in[(o * len + j) * inner + i]);
        a[j].key = descending ? (uint32_t)~k : k;
        a[j].pos = j;
      }
      merge_sort_kv(a, tmp, len);
      for (int64_t r = 0; r < len; ++r) {
        out_index[(o * len + r) * inner + i] = a[r].pos;
        out_values[(o * len + r) * inner + i] = in[(o * len + a[r].pos) * inner + i];
      }
    }
  free(a);
  free(tmp);
}

/* Stable argsort of int64 keys (treated as unsigned): restates the
 * key.argsort() under torch_sparse.coalesce and the dst-sort of the plan. */
void oracle_argsort_u64(const uint64_t* keys, int64_t n, int64_t* perm) {
  kv_t* a = (kv_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(kv_t));
  kv_t* tmp = (kv_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(kv_t));
  for (int64_t i = 0; i < n; ++i) { a[i].key = keys[i]; a[i].pos = i; }
  merge_sort_kv(a, tmp, n);
  for (int64_t i = 0; i < n; ++i) perm[i] = a[i].pos;
  free(a);
  free(tmp);
}

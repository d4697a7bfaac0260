// pybind11 bindings: tensor checks + raw-pointer extraction; all math lives in the .cu files.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include "kernels.h"
#include "symm.h"

namespace py = pybind11;
using at::Tensor;

static int dt_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return nxd::kF32;
    case at::kBFloat16: return nxd::kBF16;
    case at::kHalf: return nxd::kF16;
    default: TORCH_CHECK(false, "unsupported dtype ", t.scalar_type());
  }
}
static cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }
#define CHECK_IN(t) TORCH_CHECK((t).is_cuda() && (t).is_contiguous(), #t " must be a contiguous CUDA tensor")

static std::vector<Tensor> rmsnorm_fwd(const Tensor& x, const Tensor& w, double eps) {
  CHECK_IN(x); CHECK_IN(w);
  TORCH_CHECK(x.dim() == 2 && w.numel() == x.size(1) && w.scalar_type() == x.scalar_type());
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  auto rstd = at::empty({x.size(0), 1}, x.options().dtype(at::kFloat));
  if (x.size(0) > 0)
    nxd::rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(), (int)x.size(0), (int)x.size(1),
                     (float)eps, dt_code(x), stream());
  return {y, rstd};
}

static std::vector<Tensor> rmsnorm_bwd(const Tensor& g, const Tensor& x, const Tensor& w, const Tensor& rstd) {
  CHECK_IN(g); CHECK_IN(x); CHECK_IN(w); CHECK_IN(rstd);
  c10::cuda::CUDAGuard guard(x.device());
  const int rows = (int)x.size(0), H = (int)x.size(1);
  auto dx = at::empty_like(x);
  auto dw = at::zeros({H}, x.options().dtype(at::kFloat));
  if (rows > 0) {
    auto partial = at::empty({nxd::rmsnorm_bwd_num_partials(rows), H}, x.options().dtype(at::kFloat));
    nxd::rmsnorm_bwd(g.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), dx.data_ptr(),
                     partial.data_ptr<float>(), dw.data_ptr<float>(), rows, H, dt_code(x), stream());
  }
  return {dx, dw};
}

static std::vector<Tensor> add_rmsnorm_fwd(const Tensor& x, const Tensor& r, const Tensor& w, double eps) {
  CHECK_IN(x); CHECK_IN(r); CHECK_IN(w);
  TORCH_CHECK(x.dim() == 2 && r.sizes() == x.sizes() && w.numel() == x.size(1) && w.scalar_type() == x.scalar_type() &&
              r.scalar_type() == x.scalar_type());
  c10::cuda::CUDAGuard g(x.device());
  auto h = at::empty_like(x);
  auto y = at::empty_like(x);
  auto rstd = at::empty({x.size(0), 1}, x.options().dtype(at::kFloat));
  if (x.size(0) > 0)
    nxd::add_rmsnorm_fwd(x.data_ptr(), r.data_ptr(), w.data_ptr(), h.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(),
                         (int)x.size(0), (int)x.size(1), (float)eps, dt_code(x), stream());
  return {y, h, rstd};
}

static std::vector<Tensor> add_rmsnorm_bwd(const Tensor& gy, const Tensor& gh, const Tensor& h, const Tensor& w, const Tensor& rstd) {
  CHECK_IN(gy); CHECK_IN(gh); CHECK_IN(h); CHECK_IN(w); CHECK_IN(rstd);
  c10::cuda::CUDAGuard guard(h.device());
  const int rows = (int)h.size(0), H = (int)h.size(1);
  auto dh = at::empty_like(h);
  auto dw = at::zeros({H}, h.options().dtype(at::kFloat));
  if (rows > 0) {
    auto partial = at::empty({nxd::rmsnorm_bwd_num_partials(rows), H}, h.options().dtype(at::kFloat));
    nxd::add_rmsnorm_bwd(gy.data_ptr(), gh.data_ptr(), h.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), dh.data_ptr(),
                         partial.data_ptr<float>(), dw.data_ptr<float>(), rows, H, dt_code(h), stream());
  }
  return {dh, dw};
}

static Tensor swiglu_fwd(const Tensor& gu) {
  CHECK_IN(gu);
  c10::cuda::CUDAGuard g(gu.device());
  const int I = (int)gu.size(1) / 2;
  auto out = at::empty({gu.size(0), I}, gu.options());
  if (gu.numel()) nxd::swiglu_fwd(gu.data_ptr(), out.data_ptr(), gu.size(0), I, dt_code(gu), stream());
  return out;
}
static Tensor swiglu_bwd(const Tensor& go, const Tensor& gu) {
  CHECK_IN(go); CHECK_IN(gu);
  c10::cuda::CUDAGuard g(gu.device());
  auto d = at::empty_like(gu);
  if (gu.numel()) nxd::swiglu_bwd(go.data_ptr(), gu.data_ptr(), d.data_ptr(), gu.size(0), (int)gu.size(1) / 2, dt_code(gu), stream());
  return d;
}

static Tensor rope_apply(const Tensor& x, const Tensor& cos_t, const Tensor& sin_t, double sign) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.stride(3) == 1, "rope: x must be [B,S,H,D] with contiguous D");
  CHECK_IN(cos_t); CHECK_IN(sin_t);
  TORCH_CHECK(cos_t.scalar_type() == at::kFloat && cos_t.size(0) >= x.size(1) && cos_t.size(1) == x.size(3) / 2);
  c10::cuda::CUDAGuard g(x.device());
  auto out = at::empty(x.sizes(), x.options());
  const int es = (int)x.element_size();
  (void)es;      // unaligned / odd rotary dims take the scalar kernel (csrc/elementwise.cu rope_scalar_kernel)
  if (x.numel())
    nxd::rope_apply(x.data_ptr(), out.data_ptr(), cos_t.data_ptr<float>(), sin_t.data_ptr<float>(), (int)x.size(0),
                    (int)x.size(1), (int)x.size(2), (int)x.size(3), x.stride(0), x.stride(1), x.stride(2), (float)sign,
                    dt_code(x), stream());
  return out;
}

static Tensor ce_stats(const Tensor& logits, const Tensor& target, int64_t vocab_start) {
  CHECK_IN(logits); CHECK_IN(target);
  TORCH_CHECK(target.scalar_type() == at::kLong && logits.dim() == 2 && target.numel() == logits.size(0));
  c10::cuda::CUDAGuard g(logits.device());
  auto stats = at::empty({logits.size(0), 4}, logits.options().dtype(at::kFloat));
  if (logits.size(0))
    nxd::ce_stats(logits.data_ptr(), target.data_ptr<int64_t>(), stats.data_ptr<float>(), (int)logits.size(0),
                  (int)logits.size(1), (int)vocab_start, dt_code(logits), stream());
  return stats;
}
static Tensor ce_backward(const Tensor& logits, const Tensor& target, const Tensor& lse, const Tensor& gout,
                          int64_t vocab_start, double smoothing, int64_t vocab) {
  CHECK_IN(logits); CHECK_IN(target); CHECK_IN(lse); CHECK_IN(gout);
  c10::cuda::CUDAGuard g(logits.device());
  auto grad = at::empty_like(logits);
  if (logits.size(0))
    nxd::ce_backward(logits.data_ptr(), target.data_ptr<int64_t>(), lse.data_ptr<float>(), gout.data_ptr<float>(),
                     grad.data_ptr(), (int)logits.size(0), (int)logits.size(1), (int)vocab_start, (float)smoothing,
                     (int)vocab, dt_code(logits), stream());
  return grad;
}

static std::vector<nxd::TensorRef> refs(const std::vector<Tensor>& ts) {
  std::vector<nxd::TensorRef> r;
  r.reserve(ts.size());
  for (auto& t : ts) {
    TORCH_CHECK(t.is_cuda() && t.is_contiguous());
    r.push_back({t.data_ptr(), (long)t.numel()});
  }
  return r;
}
static void multi_tensor_sq_norm(const std::vector<Tensor>& ts, Tensor out) {
  if (ts.empty()) return;
  c10::cuda::CUDAGuard g(ts[0].device());
  nxd::multi_tensor_sq_norm(refs(ts), dt_code(ts[0]), out.data_ptr<float>(), stream());
}
static void multi_tensor_scale(const std::vector<Tensor>& ts, const Tensor& scale) {
  if (ts.empty()) return;
  c10::cuda::CUDAGuard g(ts[0].device());
  nxd::multi_tensor_scale(refs(ts), dt_code(ts[0]), scale.data_ptr<float>(), stream());
}
static void fused_adamw(const std::vector<Tensor>& p, const std::vector<Tensor>& g, const std::vector<Tensor>& m,
                        const std::vector<Tensor>& v, const std::vector<Tensor>& lowp, double lr, double b1, double b2,
                        double eps, double wd, double bc1, double bc2, const Tensor& grad_scale, bool hf_form) {
  if (p.empty()) return;
  TORCH_CHECK(p[0].scalar_type() == at::kFloat && m[0].scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(p[0].device());
  nxd::fused_adamw(refs(p), refs(g), refs(m), refs(v), refs(lowp), dt_code(g[0]), lowp.empty() ? 0 : dt_code(lowp[0]),
                   (float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (float)bc1, (float)bc2,
                   grad_scale.data_ptr<float>(), hf_form ? 1 : 0, stream());
}

static void gemm_fp8(const Tensor& a, const Tensor& b, Tensor out, const Tensor& scale_a, const Tensor& scale_b) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kFloat8_e4m3fn && b.scalar_type() == at::kFloat8_e4m3fn);
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && out.is_contiguous() && out.scalar_type() == at::kBFloat16);
  const int M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K && out.size(0) == M && out.size(1) == N && K % 16 == 0 && N % 8 == 0);
  TORCH_CHECK(scale_a.scalar_type() == at::kFloat && scale_a.numel() == M && scale_b.scalar_type() == at::kFloat &&
              scale_b.numel() == N && scale_a.is_contiguous() && scale_b.is_contiguous());
  c10::cuda::CUDAGuard guard(a.device());
  nxd::gemm_fp8(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, scale_a.data_ptr<float>(), scale_b.data_ptr<float>(),
                stream());
}

static Tensor oneshot_allreduce(const Tensor& x, const Tensor& peer_bufs, const Tensor& peer_flags, int64_t slot_bytes,
                                Tensor state, int64_t rank, int64_t world) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && (x.scalar_type() == at::kBFloat16 || x.scalar_type() == at::kFloat));
  const long nbytes = x.numel() * x.element_size();
  TORCH_CHECK(nbytes % 16 == 0 && nbytes <= slot_bytes && state.scalar_type() == at::kInt && state.numel() >= 2);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  int ctas = (int)std::max<long>(1, std::min<long>(64, nbytes / 16384));
  nxd::oneshot_allreduce(x.data_ptr(), out.data_ptr(), peer_bufs.data_ptr<int64_t>(), peer_flags.data_ptr<int64_t>(), slot_bytes,
                         (uint32_t*)state.data_ptr(), (int)rank, (int)world, x.numel(), dt_code(x), ctas, stream());
  return out;
}

// ---- row selection (sampling) -------------------------------------------------------------------------
static py::tuple row_argmax(const Tensor& x, int64_t index_offset) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1, "row_argmax: [rows, V] with contiguous last dim");
  c10::cuda::CUDAGuard g(x.device());
  auto val = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  auto idx = at::empty({x.size(0)}, x.options().dtype(at::kLong));
  nxd::row_argmax(x.data_ptr(), val.data_ptr<float>(), idx.data_ptr<long>(), (int)x.size(0), (int)x.size(1), x.stride(0), index_offset,
                  dt_code(x), stream());
  return py::make_tuple(val, idx);
}
static bool row_topk_supported(int64_t V, int64_t k) { return nxd::row_topk_supported((int)V, (int)k); }
static py::tuple row_topk(const Tensor& x, int64_t k, int64_t index_offset) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1, "row_topk: [rows, V] with contiguous last dim");
  c10::cuda::CUDAGuard g(x.device());
  auto vals = at::empty({x.size(0), k}, x.options().dtype(at::kFloat));
  auto idx = at::empty({x.size(0), k}, x.options().dtype(at::kLong));
  nxd::row_topk(x.data_ptr(), vals.data_ptr<float>(), idx.data_ptr<long>(), (int)x.size(0), (int)x.size(1), (int)k, x.stride(0),
                index_offset, dt_code(x), stream());
  return py::make_tuple(vals, idx);
}

// expert_index [T, k] (int32 / int64) → (block_to_expert [nb], token_position_to_id [nb*B], tokens_per_expert [E]) int64
static py::tuple moe_block_metadata(const Tensor& expert_index, int64_t num_experts, int64_t block_size, int64_t num_blocks) {
  CHECK_IN(expert_index);
  TORCH_CHECK(expert_index.dim() == 2 && (expert_index.scalar_type() == at::kLong || expert_index.scalar_type() == at::kInt));
  c10::cuda::CUDAGuard g(expert_index.device());
  auto o = expert_index.options().dtype(at::kLong);
  auto b2e = at::empty({num_blocks}, o), tp2id = at::empty({num_blocks * block_size}, o), counts = at::empty({num_experts}, o);
  nxd::moe_block_metadata(expert_index.data_ptr(), expert_index.scalar_type() == at::kLong, expert_index.numel(), (int)expert_index.size(1),
                          (int)num_experts, (int)block_size, (int)num_blocks, b2e.data_ptr<long>(), tp2id.data_ptr<long>(),
                          counts.data_ptr<long>(), stream());
  return py::make_tuple(b2e, tp2id, counts);
}

// q [B,1,H,D], k/v [B,1,Hkv,D] (views of the fused QKV GEMV output), cache [B,L,Hkv,D]: rotate q → new tensor, rotate k and
// copy v straight into the cache row positions[b]
static Tensor decode_rope_kv(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& positions, const Tensor& cos_t,
                             const Tensor& sin_t, Tensor kc, Tensor vc) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && q.dim() == 4 && q.size(1) == 1 && q.stride(3) == 1);
  TORCH_CHECK(k.scalar_type() == at::kBFloat16 && k.dim() == 4 && k.size(1) == 1 && k.stride(3) == 1 && v.stride(3) == 1 && v.sizes() == k.sizes());
  TORCH_CHECK(kc.scalar_type() == at::kBFloat16 && kc.dim() == 4 && kc.stride(3) == 1 && vc.sizes() == kc.sizes() && vc.strides() == kc.strides());
  TORCH_CHECK(positions.scalar_type() == at::kLong && positions.is_contiguous() && positions.numel() == q.size(0));
  CHECK_IN(cos_t); CHECK_IN(sin_t);
  TORCH_CHECK(cos_t.scalar_type() == at::kFloat && cos_t.size(1) == q.size(3) / 2 && cos_t.size(0) >= kc.size(1));
  c10::cuda::CUDAGuard g(q.device());
  const int B = (int)q.size(0), H = (int)q.size(2), Hkv = (int)k.size(2), D = (int)q.size(3), L = (int)kc.size(1);
  Tensor q_out = at::empty({B, 1, H, D}, q.options());
  nxd::decode_rope_kv(q.data_ptr(), k.data_ptr(), v.data_ptr(), positions.data_ptr<long>(), cos_t.data_ptr<float>(),
                      sin_t.data_ptr<float>(), q_out.data_ptr(), kc.data_ptr(), vc.data_ptr(), B, H, Hkv, D, L, q.stride(0), q.stride(2),
                      k.stride(0), k.stride(2), v.stride(0), v.stride(2), kc.stride(0), kc.stride(1), kc.stride(2), stream());
  return q_out;
}

// ---- decode ---------------------------------------------------------------------------------------
// q [B,1,H,128]; k/v cache [B,L,Hkv,128] (any b/s/h strides, head_dim contiguous); positions [B] int64 (attend to ≤ pos)
static Tensor decode_attention(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& positions, double scale) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && q.dim() == 4 && q.size(1) == 1 && q.size(3) == 128 && q.stride(3) == 1);
  TORCH_CHECK(k.scalar_type() == at::kBFloat16 && k.dim() == 4 && k.size(3) == 128 && k.stride(3) == 1 && v.stride(3) == 1 &&
              v.sizes() == k.sizes());
  TORCH_CHECK(positions.scalar_type() == at::kLong && positions.is_cuda() && positions.is_contiguous() && positions.numel() == q.size(0));
  const int B = q.size(0), H = q.size(2), L = k.size(1), Hkv = k.size(2);
  TORCH_CHECK(H % Hkv == 0 && k.size(0) == B);
  TORCH_CHECK(q.stride(2) % 8 == 0 && k.stride(1) % 8 == 0 && k.stride(2) % 8 == 0 && v.stride(1) % 8 == 0 && v.stride(2) % 8 == 0);
  c10::cuda::CUDAGuard guard(q.device());
  // enough CTAs for ~6 per SM (the kernel is latency-bound per CTA: 8 cache rows in flight), at least 64 rows per split
  int splits = (6 * 148 + B * Hkv - 1) / (B * Hkv);
  splits = std::max(1, std::min(std::min(splits, 64), (L + 63) / 64));
  Tensor out = at::empty({B, 1, H, 128}, q.options());
  Tensor part_o = at::empty({B * H * splits, 128}, q.options().dtype(at::kFloat));
  Tensor part_ml = at::empty({B * H * splits, 2}, q.options().dtype(at::kFloat));
  const long ks[3] = {k.stride(0), k.stride(1), k.stride(2)}, vs[3] = {v.stride(0), v.stride(1), v.stride(2)};
  nxd::decode_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), positions.data_ptr<long>(), out.data_ptr(),
                        part_o.data_ptr<float>(), part_ml.data_ptr<float>(), B, H, Hkv, L, ks, vs, q.stride(0), q.stride(2),
                        out.stride(0), out.stride(2), (float)scale, splits, stream());
  return out;
}
// This rank's contribution to distributed flash-decoding: un-normalised output (relative to the row maximum) [B,H,128] fp32 and
// (row maximum in natural-log units, row sum) [B,H,2] fp32 over the local cache shard; ``positions`` are LOCAL (may be < 0 or
// >= L: nothing / everything of this shard is visible).
static std::vector<Tensor> decode_attention_partial(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& positions,
                                                    double scale) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && q.dim() == 4 && q.size(1) == 1 && q.size(3) == 128 && q.stride(3) == 1);
  TORCH_CHECK(k.scalar_type() == at::kBFloat16 && k.dim() == 4 && k.size(3) == 128 && k.stride(3) == 1 && v.stride(3) == 1 &&
              v.sizes() == k.sizes());
  TORCH_CHECK(positions.scalar_type() == at::kLong && positions.is_cuda() && positions.is_contiguous() && positions.numel() == q.size(0));
  const int B = q.size(0), H = q.size(2), L = k.size(1), Hkv = k.size(2);
  TORCH_CHECK(H % Hkv == 0 && k.size(0) == B);
  TORCH_CHECK(q.stride(2) % 8 == 0 && k.stride(1) % 8 == 0 && k.stride(2) % 8 == 0 && v.stride(1) % 8 == 0 && v.stride(2) % 8 == 0);
  c10::cuda::CUDAGuard guard(q.device());
  int splits = (6 * 148 + B * Hkv - 1) / (B * Hkv);
  splits = std::max(1, std::min(std::min(splits, 64), (L + 63) / 64));
  Tensor fin_o = at::empty({B, H, 128}, q.options().dtype(at::kFloat));
  Tensor fin_ml = at::empty({B, H, 2}, q.options().dtype(at::kFloat));
  Tensor part_o = at::empty({B * H * splits, 128}, q.options().dtype(at::kFloat));
  Tensor part_ml = at::empty({B * H * splits, 2}, q.options().dtype(at::kFloat));
  const long ks[3] = {k.stride(0), k.stride(1), k.stride(2)}, vs[3] = {v.stride(0), v.stride(1), v.stride(2)};
  nxd::decode_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), positions.data_ptr<long>(), nullptr, part_o.data_ptr<float>(),
                        part_ml.data_ptr<float>(), B, H, Hkv, L, ks, vs, q.stride(0), q.stride(2), 0, 0, (float)scale, splits,
                        stream(), fin_o.data_ptr<float>(), fin_ml.data_ptr<float>());
  return {fin_o, fin_ml};
}
// x [T,H] bf16; gamma [H] bf16 or none; router_w [E,H] bf16; router_bias [E] fp32 or none; w_gu [El,H,2I]; w_dn [El,I,H]
// → (out [T,H] bf16 partial over this rank's experts / intermediate shard, router logits [T,E] fp32, top-k idx [T,K], top-k w [T,K])
static std::vector<Tensor> moe_block_tkg(const Tensor& x, const c10::optional<Tensor>& gamma, const Tensor& router_w,
                                         const c10::optional<Tensor>& router_bias, const Tensor& w_gu, const Tensor& w_dn,
                                         int64_t e0, int64_t top_k, double eps, int64_t router_act, bool act_over_topk, bool normalize,
                                         bool pre_scale, bool round_logits, int64_t act, double act_alpha, double act_beta,
                                         double gate_lo, double gate_hi, double up_lo, double up_hi, bool cooperative) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous());
  TORCH_CHECK(router_w.scalar_type() == at::kBFloat16 && router_w.dim() == 2 && router_w.is_contiguous() && router_w.size(1) == x.size(1));
  TORCH_CHECK(w_gu.scalar_type() == at::kBFloat16 && w_gu.dim() == 3 && w_gu.is_contiguous() && w_gu.size(1) == x.size(1) &&
              w_gu.size(2) % 2 == 0);
  TORCH_CHECK(w_dn.scalar_type() == at::kBFloat16 && w_dn.dim() == 3 && w_dn.is_contiguous() && w_dn.size(0) == w_gu.size(0) &&
              w_dn.size(1) * 2 == w_gu.size(2) && w_dn.size(2) == x.size(1));
  const int T = x.size(0), H = x.size(1), E = router_w.size(0), El = w_gu.size(0), I = w_dn.size(1), K = top_k;
  TORCH_CHECK(nxd::moe_block_tkg_supported(T, H, E, I, K), "moe_block_tkg: unsupported shape");
  TORCH_CHECK(e0 >= 0 && e0 + El <= E);
  const void* g = nullptr;
  if (gamma.has_value()) {
    TORCH_CHECK(gamma->scalar_type() == at::kBFloat16 && gamma->is_contiguous() && gamma->numel() == H);
    g = gamma->data_ptr();
  }
  const float* rb = nullptr;
  if (router_bias.has_value()) {
    TORCH_CHECK(router_bias->scalar_type() == at::kFloat && router_bias->is_contiguous() && router_bias->numel() == E);
    rb = router_bias->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(x.device());
  auto f32 = x.options().dtype(at::kFloat);
  Tensor out = at::empty({T, H}, x.options());
  Tensor logits = at::empty({T, E}, f32);
  Tensor idx = at::empty({T, K}, x.options().dtype(at::kLong));
  Tensor w = at::empty({T, K}, f32);
  Tensor gu = at::empty({(long)T * K, 2L * I}, f32);
  Tensor yacc = at::empty({T, H}, f32);
  Tensor bar = at::empty({1}, x.options().dtype(at::kInt));
  nxd::moe_block_tkg(x.data_ptr(), g, router_w.data_ptr(), rb, w_gu.data_ptr(), w_dn.data_ptr(), logits.data_ptr<float>(),
                     gu.data_ptr<float>(), yacc.data_ptr<float>(), out.data_ptr(), idx.data_ptr<long>(), w.data_ptr<float>(),
                     (unsigned*)bar.data_ptr<int>(), T, H, E, El, (int)e0, I, K, (float)eps, (int)router_act, act_over_topk, normalize,
                     pre_scale, round_logits, (int)act, (float)act_alpha, (float)act_beta, (float)gate_lo, (float)gate_hi, (float)up_lo,
                     (float)up_hi, cooperative, stream());
  return {out, logits, idx, w};
}
// a [M,K], b [N,K]: fp8 bytes; sfa [ceil(M/128), K/128, 512], sfb [ceil(N/128), K/128, 512]: tiled E8M0 scales → [M,N] bf16
static Tensor gemm_mxfp8(const Tensor& a, const Tensor& b, const Tensor& sfa, const Tensor& sfb, int64_t a_fmt, int64_t b_fmt) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(sfa); CHECK_IN(sfb);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.element_size() == 1 && b.element_size() == 1);
  // byte tensors: K bytes per row for fp8 formats, K/2 for packed e2m1 (format 5)
  const int M = a.size(0), N = b.size(0), K = a_fmt == 5 ? 2 * a.size(1) : a.size(1);
  TORCH_CHECK((b_fmt == 5 ? 2 * b.size(1) : b.size(1)) == K, "gemm_mxfp8: K of a and b differ");
  TORCH_CHECK(K % 128 == 0 && N % 8 == 0, "gemm_mxfp8: K % 128 == 0 and N % 8 == 0");
  TORCH_CHECK(sfa.scalar_type() == at::kByte && sfb.scalar_type() == at::kByte);
  TORCH_CHECK(sfa.numel() == (long)((M + 127) / 128) * (K / 128) * 512 && sfb.numel() == (long)((N + 127) / 128) * (K / 128) * 512,
              "gemm_mxfp8: scale tensors must be tiled [rows/128, K/128, 512]");
  c10::cuda::CUDAGuard guard(a.device());
  Tensor out = at::empty({M, N}, a.options().dtype(at::kBFloat16));
  nxd::gemm_mxfp8(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), out.data_ptr(), M, N, K, (int)a_fmt, (int)b_fmt, stream());
  return out;
}
// a [M,K/2], b [N,K/2]: packed e2m1 bytes; sfa [ceil(M/128), K/(4·vs), 512], sfb likewise: tiled scales (E8M0 for vs 32, UE4M3 for
// vs 16) → alpha · product, [M,N] bf16
static Tensor gemm_f4(const Tensor& a, const Tensor& b, const Tensor& sfa, const Tensor& sfb, int64_t vec_size, double alpha) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(sfa); CHECK_IN(sfb);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.element_size() == 1 && b.element_size() == 1);
  const int M = a.size(0), N = b.size(0), K = 2 * a.size(1);
  TORCH_CHECK(2 * b.size(1) == K, "gemm_f4: K of a and b differ");
  TORCH_CHECK(K % 256 == 0 && N % 8 == 0, "gemm_f4: K % 256 == 0 and N % 8 == 0");
  TORCH_CHECK(vec_size == 32 || vec_size == 16, "gemm_f4: vec_size 32 (MXFP4) or 16 (NVFP4)");
  TORCH_CHECK(sfa.scalar_type() == at::kByte && sfb.scalar_type() == at::kByte);
  const long kc = K / (4 * vec_size);
  TORCH_CHECK(sfa.numel() == (long)((M + 127) / 128) * kc * 512 && sfb.numel() == (long)((N + 127) / 128) * kc * 512,
              "gemm_f4: scale tensors must be tiled [rows/128, K/(4*vec_size), 512]");
  c10::cuda::CUDAGuard guard(a.device());
  Tensor out = at::empty({M, N}, a.options().dtype(at::kBFloat16));
  nxd::gemm_f4(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), out.data_ptr(), M, N, K, (int)vec_size, (float)alpha, stream());
  return out;
}
// x [M<=8, K] bf16; w = MX byte stream of [N, K] (fp4: K/2 bytes per row, fp8: K), any integer dtype view; scale [N, K/32] uint8
static Tensor gemv_mx(const Tensor& x, const Tensor& w, const Tensor& scale, int64_t fmt, const c10::optional<Tensor>& residual) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous() && x.size(0) >= 1 && x.size(0) <= 8);
  TORCH_CHECK(w.is_cuda() && w.is_contiguous() && scale.is_cuda() && scale.is_contiguous() && scale.scalar_type() == at::kByte &&
              scale.dim() == 2);
  const int M = x.size(0), K = x.size(1), N = scale.size(0);
  TORCH_CHECK(K % 32 == 0 && scale.size(1) == K / 32, "gemv_mx: scale must be [N, K/32]");
  const long row_bytes = fmt == 0 ? K / 2 : K;
  TORCH_CHECK((long)(w.numel() * w.element_size()) == (long)N * row_bytes, "gemv_mx: weight bytes do not match [N, K] in this format");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(w.data_ptr()) % 16 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({M, N}, x.options());
  const void* res = nullptr;
  if (residual.has_value()) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == y.sizes() && residual->scalar_type() == at::kBFloat16);
    res = residual->data_ptr();
  }
  nxd::gemv_mx(x.data_ptr(), w.data_ptr(), scale.data_ptr(), res, y.data_ptr(), M, N, K, (int)fmt, stream());
  return y;
}
// x [S, K] bf16 (one row per (token, slot)); w = MX byte stream of [E, N, K]; scale [E, N, K/32]; expert [S] int64 → [S, N] bf16
static Tensor gemv_mx_grouped(const Tensor& x, const Tensor& w, const Tensor& scale, const Tensor& expert, int64_t fmt) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous());
  TORCH_CHECK(w.is_cuda() && w.is_contiguous() && scale.is_cuda() && scale.is_contiguous() && scale.scalar_type() == at::kByte &&
              scale.dim() == 3);
  TORCH_CHECK(expert.is_cuda() && expert.scalar_type() == at::kLong && expert.is_contiguous() && expert.numel() == x.size(0));
  const int S = x.size(0), K = x.size(1), E = scale.size(0), N = scale.size(1);
  TORCH_CHECK(K % 32 == 0 && scale.size(2) == K / 32, "gemv_mx_grouped: scale must be [E, N, K/32]");
  const long row_bytes = fmt == 0 ? K / 2 : K;
  TORCH_CHECK((long)(w.numel() * w.element_size()) == (long)E * N * row_bytes, "gemv_mx_grouped: weight bytes do not match [E, N, K]");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({S, N}, x.options());
  nxd::gemv_mx_grouped(x.data_ptr(), w.data_ptr(), scale.data_ptr(), expert.data_ptr<long>(), y.data_ptr(), S, N, K, E, (int)fmt, stream());
  return y;
}
static Tensor gemv(const Tensor& x, const Tensor& w, const c10::optional<Tensor>& residual) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x.dim() == 2 && w.dim() == 2);
  TORCH_CHECK(x.is_contiguous() && w.is_contiguous() && x.size(1) == w.size(1) && x.size(0) >= 1 && x.size(0) <= 8 && x.size(1) % 8 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({x.size(0), w.size(0)}, x.options());
  const void* res = nullptr;
  if (residual.has_value()) {
    TORCH_CHECK(residual->is_contiguous() && residual->sizes() == y.sizes() && residual->scalar_type() == at::kBFloat16);
    res = residual->data_ptr();
  }
  nxd::gemv_bf16(x.data_ptr(), w.data_ptr(), res, y.data_ptr(), x.size(0), w.size(0), x.size(1), stream());
  return y;
}

// ---- grouped (MoE) GEMMs -------------------------------------------------------------------------
static Tensor grouped_gemm(const Tensor& a, const Tensor& w, const Tensor& block_expert, int64_t block_rows, bool trans_b) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && a.dim() == 2 && a.is_contiguous());
  TORCH_CHECK(w.scalar_type() == at::kBFloat16 && w.dim() == 3 && w.is_contiguous());
  TORCH_CHECK(block_expert.scalar_type() == at::kInt && block_expert.is_cuda());
  const int M = a.size(0), K = a.size(1), E = w.size(0);
  const int N = trans_b ? w.size(1) : w.size(2);
  TORCH_CHECK((trans_b ? w.size(2) : w.size(1)) == K && M % block_rows == 0 && K % 8 == 0 && N % 8 == 0);
  TORCH_CHECK(block_expert.numel() * block_rows == M);
  c10::cuda::CUDAGuard guard(a.device());
  Tensor out = at::empty({M, N}, a.options());
  nxd::grouped_gemm_bf16(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, E, trans_b, block_expert.data_ptr<int>(),
                         (int)block_rows, nxd::kBF16, stream());
  return out;
}
// dW[e] = x[rows of e]ᵀ · dy[rows of e]  → out [E, Mo, No] (bf16 or fp32), optionally accumulated
static void grouped_wgrad(const Tensor& x, const Tensor& dy, Tensor out, const Tensor& seg_first_block, int64_t block_rows,
                          bool accumulate) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.is_contiguous() && dy.scalar_type() == at::kBFloat16 &&
              dy.is_contiguous() && out.is_contiguous() && out.dim() == 3);
  TORCH_CHECK(seg_first_block.scalar_type() == at::kInt && seg_first_block.numel() == out.size(0) + 1);
  TORCH_CHECK(x.size(0) == dy.size(0) && out.size(1) == x.size(1) && out.size(2) == dy.size(1));
  TORCH_CHECK(x.size(1) % 8 == 0 && dy.size(1) % 8 == 0);
  c10::cuda::CUDAGuard guard(x.device());
  nxd::grouped_wgrad_bf16(x.data_ptr(), dy.data_ptr(), out.data_ptr(), x.size(0), x.size(1), dy.size(1), out.size(0),
                          seg_first_block.data_ptr<int>(), (int)block_rows, dt_code(out), accumulate, stream());
}

// ---- attention ----------------------------------------------------------------------------------
static void check_qkv(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16 && t.dim() == 4 && t.size(3) == 128 && t.stride(3) == 1,
              name, ": expected a bf16 CUDA [B,S,H,128] view with contiguous head_dim");
}
// returns (out [B,S,H,D] view — memory laid out [S,B,H,D] when sbhd_out —, lse [B,H,S] fp32)
static std::vector<Tensor> flash_attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, bool causal, double scale,
                                          bool sbhd_out) {
  check_qkv(q, "q"); check_qkv(k, "k"); check_qkv(v, "v");
  const int B = q.size(0), S = q.size(1), H = q.size(2), Skv = k.size(1), Hkv = k.size(2);
  TORCH_CHECK(k.size(0) == B && v.size(0) == B && v.size(1) == Skv && v.size(2) == Hkv && H % Hkv == 0);
  c10::cuda::CUDAGuard guard(q.device());
  Tensor out = sbhd_out ? at::empty({S, B, H, 128}, q.options()).transpose(0, 1) : at::empty({B, S, H, 128}, q.options());
  Tensor lse = at::empty({B, H, S}, q.options().dtype(at::kFloat));
  const long qs[3] = {q.stride(0), q.stride(1), q.stride(2)}, ks[3] = {k.stride(0), k.stride(1), k.stride(2)};
  const long vs[3] = {v.stride(0), v.stride(1), v.stride(2)}, os[3] = {out.stride(0), out.stride(1), out.stride(2)};
  nxd::flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), B, S, Skv, H, Hkv,
                      qs, ks, vs, os, (float)scale, causal, stream());
  return {out, lse};
}

static std::vector<Tensor> flash_attn_bwd(const Tensor& go, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& o,
                                          const Tensor& lse, bool causal, double scale, bool sbhd_out,
                                          const c10::optional<Tensor>& dlse) {
  check_qkv(q, "q"); check_qkv(k, "k"); check_qkv(v, "v"); check_qkv(go, "grad_out"); check_qkv(o, "out");
  const int B = q.size(0), S = q.size(1), H = q.size(2), Skv = k.size(1), Hkv = k.size(2);
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == (int64_t)B * H * S);
  c10::cuda::CUDAGuard guard(q.device());
  auto mk = [&](int s, int h) {
    return sbhd_out ? at::empty({s, B, h, 128}, q.options()).transpose(0, 1) : at::empty({B, s, h, 128}, q.options());
  };
  Tensor dq = mk(S, H), dk = mk(Skv, Hkv), dv = mk(Skv, Hkv);
  const int S_pad = (S + 63) / 64 * 64;
  Tensor stats = at::empty({2, (int64_t)B * H * S_pad + 64}, q.options().dtype(at::kFloat));
  Tensor dq_acc = at::empty({B, H, S_pad, 128}, q.options().dtype(at::kFloat));
  auto st3 = [](const Tensor& t, long* out) { out[0] = t.stride(0); out[1] = t.stride(1); out[2] = t.stride(2); };
  long gs[3], qs[3], ks[3], vs[3], os[3], dqs[3], dks[3], dvs[3];
  st3(go, gs); st3(q, qs); st3(k, ks); st3(v, vs); st3(o, os); st3(dq, dqs); st3(dk, dks); st3(dv, dvs);
  const float* dlse_p = nullptr;
  if (dlse.has_value()) {
    TORCH_CHECK(dlse->scalar_type() == at::kFloat && dlse->is_contiguous() && dlse->numel() == lse.numel());
    dlse_p = dlse->data_ptr<float>();
  }
  nxd::flash_attn_bwd(go.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr<float>(), dlse_p,
                      dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), stats.data_ptr<float>(), dq_acc.data_ptr<float>(), B, S,
                      Skv, H, Hkv, S_pad, gs, qs, ks, vs, os, dqs, dks, dvs, (float)scale, causal, stream());
  return {dq, dk, dv};
}

// ---- GEMM ---------------------------------------------------------------------------------------
static void gemm_bf16(const Tensor& a, const Tensor& b, Tensor out, bool trans_a, bool trans_b, bool accumulate) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(out);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && a.dim() == 2 && b.dim() == 2);
  const int M = (int)(trans_a ? a.size(1) : a.size(0)), K = (int)(trans_a ? a.size(0) : a.size(1));
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  TORCH_CHECK((trans_b ? b.size(1) : b.size(0)) == K && out.size(0) == M && out.size(1) == N, "gemm shape mismatch");
  c10::cuda::CUDAGuard g(a.device());
  nxd::GemmComm none;
  nxd::gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, trans_a, trans_b, dt_code(out), accumulate, none,
                 nullptr, stream());
}

static void gemm_bf16_2cta(const Tensor& a, const Tensor& b, Tensor out, bool trans_a, bool trans_b, bool accumulate) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(out);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && a.dim() == 2 && b.dim() == 2);
  const int M = (int)(trans_a ? a.size(1) : a.size(0)), K = (int)(trans_a ? a.size(0) : a.size(1));
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  TORCH_CHECK((trans_b ? b.size(1) : b.size(0)) == K && out.size(0) == M && out.size(1) == N, "gemm shape mismatch");
  c10::cuda::CUDAGuard g(a.device());
  nxd::gemm_bf16_2cta(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, trans_a, trans_b, dt_code(out), accumulate, stream());
}

// all-gather(A shards along rows) → GEMM.  `a_shard` [M/world, K]; gathered A lives in the symmetric payload at
// buf_offset (every rank), out = A_full @ op(B).
static void ag_gemm_bf16(const Tensor& a_shard, const Tensor& b, Tensor out, bool trans_b, int64_t local_buf_ptr,
                         const Tensor& peer_bufs, const Tensor& peer_flags, int64_t buf_offset, int64_t flag_offset,
                         int64_t epoch, int64_t rank, int64_t world, int64_t comm_sms) {
  CHECK_IN(a_shard); CHECK_IN(b); CHECK_IN(out);
  const int Ms = (int)a_shard.size(0), K = (int)a_shard.size(1), M = Ms * (int)world;
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  TORCH_CHECK(out.size(0) == M && out.size(1) == N);
  c10::cuda::CUDAGuard g(a_shard.device());
  nxd::GemmComm c;
  c.mode = 1; c.rank = (int)rank; c.world = (int)world;
  c.peer_bufs = peer_bufs.data_ptr<int64_t>(); c.peer_flags = peer_flags.data_ptr<int64_t>();
  c.buf_offset = buf_offset; c.flag_offset = (int)flag_offset; c.epoch = (uint32_t)epoch; c.comm_sms = (int)comm_sms;
  // A operand = this rank's gathered buffer
  const void* a_full = (const void*)(local_buf_ptr + buf_offset);
  nxd::gemm_bf16(a_full, b.data_ptr(), out.data_ptr(), M, N, K, false, trans_b, dt_code(out), false, c,
                 a_shard.data_ptr(), stream());
}

// GEMM → reduce-scatter over rows.  a [M, K]; out [M/world, N] bf16 (sum over ranks of rows owned by this rank).
static void gemm_rs_bf16(const Tensor& a, const Tensor& b, Tensor out, bool trans_b, int64_t local_buf_ptr,
                         const Tensor& peer_bufs, const Tensor& peer_flags, int64_t buf_offset, int64_t flag_offset,
                         const std::vector<int64_t>& targets, int64_t rank, int64_t world) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(out);
  const int M = (int)a.size(0), K = (int)a.size(1);
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  TORCH_CHECK(out.size(0) * world == M && out.size(1) == N && out.scalar_type() == at::kBFloat16);
  (void)local_buf_ptr;
  c10::cuda::CUDAGuard g(a.device());
  nxd::GemmComm c;
  c.mode = 2; c.rank = (int)rank; c.world = (int)world;
  c.peer_bufs = peer_bufs.data_ptr<int64_t>(); c.peer_flags = peer_flags.data_ptr<int64_t>();
  c.buf_offset = buf_offset; c.flag_offset = (int)flag_offset; c.epoch = 0;
  uint32_t tg[64] = {0};
  for (size_t i = 0; i < targets.size() && i < 64; ++i) tg[i] = (uint32_t)targets[i];
  c.rs_targets = tg;
  nxd::gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, false, trans_b, nxd::kBF16, false, c, nullptr,
                 stream());
}

// CTA-pair fused TP kernels.  mode 1: a = shard [M/world,K] → out [M,N];  mode 2: a = [M,K] → out [M/world,N]
static void tp_gemm_2cta(int64_t mode, const Tensor& a, const Tensor& b, Tensor out, Tensor partial, bool trans_b,
                         int64_t local_buf_ptr, const Tensor& peer_bufs, const Tensor& peer_flags, int64_t buf_offset,
                         int64_t flag_offset, int64_t epoch, int64_t rank, int64_t world, int64_t comm_ctas,
                         Tensor counters, int64_t gemm_done_target) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(out);
  c10::cuda::CUDAGuard g(a.device());
  const int K = (int)a.size(1);
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  auto* ctr = (uint32_t*)counters.data_ptr();           // [0] = gemm_done, [64..] = tile_done per 128-row block
  if (mode == 1) {
    const int M = (int)a.size(0) * (int)world;
    TORCH_CHECK(out.size(0) == M && out.size(1) == N);
    nxd::gemm_bf16_2cta_tp(1, (const void*)(local_buf_ptr + buf_offset), b.data_ptr(), out.data_ptr(), nullptr, a.data_ptr(),
                           M, N, K, trans_b, (int)rank, (int)world, peer_bufs.data_ptr<int64_t>(),
                           peer_flags.data_ptr<int64_t>(), buf_offset, (int)flag_offset, (uint32_t)epoch, (int)comm_ctas,
                           ctr + 64, ctr, 0u, stream());
  } else {
    const int M = (int)a.size(0);
    TORCH_CHECK(out.size(0) * world == M && out.size(1) == N && partial.size(0) == M && partial.size(1) == N);
    nxd::gemm_bf16_2cta_tp(2, a.data_ptr(), b.data_ptr(), partial.data_ptr(), out.data_ptr(), nullptr, M, N, K, trans_b,
                           (int)rank, (int)world, peer_bufs.data_ptr<int64_t>(), peer_flags.data_ptr<int64_t>(), buf_offset,
                           (int)flag_offset, (uint32_t)epoch, (int)comm_ctas, ctr + 64, ctr, (uint32_t)gemm_done_target,
                           stream());
  }
}

// ---- ZeRO-1 collectives over peer memory
static void zero1_reduce_scatter(const Tensor& peer_bufs, int64_t grad_off_bytes, const Tensor& peer_flags, int64_t flag_off,
                                 int64_t epoch, int64_t rank, int64_t world, int64_t shard_numel, double scale, Tensor out,
                                 Tensor done_ctr, bool grads_fp32, int64_t sub_begin, int64_t sub_len, int64_t max_ctas) {
  TORCH_CHECK(sub_begin >= 0 && sub_len >= 0 && sub_begin + sub_len <= shard_numel && sub_begin % 8 == 0 && sub_len % 8 == 0);
  c10::cuda::CUDAGuard g(out.device());
  nxd::zero1_reduce_scatter(peer_bufs.data_ptr<int64_t>(), grad_off_bytes, peer_flags.data_ptr<int64_t>(), (int)flag_off,
                            (uint32_t)epoch, (int)rank, (int)world, shard_numel, sub_begin, sub_len, (float)scale,
                            out.data_ptr<float>(), (uint32_t*)done_ctr.data_ptr(), grads_fp32 ? nxd::kF32 : nxd::kBF16,
                            (int)max_ctas, stream());
}
static void zero1_all_gather(const Tensor& master, const Tensor& peer_bufs, int64_t param_off_bytes, const Tensor& peer_flags,
                             int64_t flag_off, int64_t epoch, int64_t rank, int64_t world, int64_t shard_numel,
                             Tensor done_ctr, bool params_bf16) {
  c10::cuda::CUDAGuard g(master.device());
  nxd::zero1_all_gather(master.data_ptr<float>(), peer_bufs.data_ptr<int64_t>(), param_off_bytes,
                        peer_flags.data_ptr<int64_t>(), (int)flag_off, (uint32_t)epoch, (int)rank, (int)world, shard_numel,
                        (uint32_t*)done_ctr.data_ptr(), params_bf16 ? nxd::kBF16 : nxd::kF32, stream());
}

// ---- symmetric memory ---------------------------------------------------------------------------
static py::tuple symm_alloc(int64_t nbytes, int64_t nflags) {
  auto a = nxd::symm_alloc((size_t)nbytes, (size_t)nflags);
  return py::make_tuple(a.id, py::bytes(a.payload_handle), py::bytes(a.flags_handle));
}
static py::tuple symm_open(int64_t id, int64_t rank, const std::vector<py::bytes>& ph, const std::vector<py::bytes>& fh) {
  std::vector<std::string> p, f;
  for (auto& b : ph) p.push_back((std::string)b);
  for (auto& b : fh) f.push_back((std::string)b);
  auto r = nxd::symm_open(id, (int)rank, p, f);
  return py::make_tuple(r.payload, r.flags);
}
static Tensor symm_view(int64_t id, int64_t offset, const std::vector<int64_t>& shape, py::object dtype) {
  size_t nbytes = 0;
  auto* base = (uint8_t*)nxd::symm_local_payload(id, &nbytes);
  auto st = torch::python::detail::py_object_to_dtype(dtype);
  int dev; cudaGetDevice(&dev);
  auto opts = at::TensorOptions().dtype(st).device(at::kCUDA, dev);
  return at::from_blob(base + offset, shape, opts);
}

// ---- NVLS symmetric memory (symm_vmm.cpp) + fused TP kernels on it (tp_nvls_sm100.cu) + stand-alone collectives ----------
static py::tuple vmm_begin(int64_t nbytes, int64_t rank, int64_t world, bool want_multicast) {
  auto b = nxd::vmm_begin((size_t)nbytes, (int)rank, (int)world, want_multicast);
  return py::make_tuple(b.id, b.sock_name, b.multicast_supported, (int64_t)b.size);
}
static py::tuple vmm_ptrs(int64_t id, bool multicast_everywhere) {
  auto p = nxd::vmm_ptrs(id, multicast_everywhere);
  return py::make_tuple(p.peer, p.multicast, (int64_t)p.size);
}
static Tensor vmm_view(int64_t id, int64_t offset, const std::vector<int64_t>& shape, py::object dtype) {
  size_t nbytes = 0;
  auto* base = (uint8_t*)nxd::vmm_local(id, &nbytes);
  auto st = torch::python::detail::py_object_to_dtype(dtype);
  int64_t need = at::elementSize(st);
  for (auto d : shape) need *= d;
  TORCH_CHECK(offset >= 0 && (size_t)(offset + need) <= nbytes, "vmm_view out of range");
  int dev; cudaGetDevice(&dev);
  auto opts = at::TensorOptions().dtype(st).device(at::kCUDA, dev);
  return at::from_blob(base + offset, shape, opts);
}

// mode 1: a = shard [M/world, K] → out [M, N]; gathered A = region payload at buf_offset.
// mode 2: a = [M, K] → out [M/world, N]; partials live in the region payload at buf_offset.  Returns reducer claims used.
static int64_t tp_gemm_nvls(int64_t mode, const Tensor& a, const Tensor& b, Tensor out, bool trans_b, const Tensor& peer_bases,
                            int64_t mc_base, int64_t local_base, int64_t buf_offset, int64_t flag_offset, int64_t epoch,
                            int64_t rank, int64_t world, int64_t comm_ctas, Tensor counters, int64_t claim_base, bool wire_fp32,
                            bool gemm_join) {
  CHECK_IN(a); CHECK_IN(b); CHECK_IN(out);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && out.scalar_type() == at::kBFloat16);
  TORCH_CHECK(counters.scalar_type() == at::kInt && counters.numel() >= 64 + 8 * 256);
  c10::cuda::CUDAGuard g(a.device());
  const int K = (int)a.size(1);
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  auto* ctr = (uint32_t*)counters.data_ptr();           // [0] = reducer claim counter, [64..] = tile_done per 128-row block
  if (mode == 1) {
    const int M = (int)a.size(0) * (int)world;
    TORCH_CHECK(out.size(0) == M && out.size(1) == N);
    return nxd::gemm_bf16_2cta_nvls(1, nullptr, b.data_ptr(), out.data_ptr(), nullptr, a.data_ptr(), M, N, K, trans_b, (int)rank,
                                    (int)world, peer_bases.data_ptr<int64_t>(), mc_base, local_base, buf_offset, flag_offset,
                                    (uint32_t)epoch, (int)comm_ctas, ctr + 64, ctr, 0u, false, true, stream());
  }
  const int M = (int)a.size(0);
  TORCH_CHECK(out.size(0) * world == M && out.size(1) == N);
  return nxd::gemm_bf16_2cta_nvls(2, a.data_ptr(), b.data_ptr(), nullptr, out.data_ptr(), nullptr, M, N, K, trans_b, (int)rank,
                                  (int)world, peer_bases.data_ptr<int64_t>(), mc_base, local_base, buf_offset, flag_offset,
                                  (uint32_t)epoch, (int)comm_ctas, ctr + 64, ctr, (uint32_t)claim_base, wire_fp32, gemm_join, stream());
}

static Tensor nvls_allreduce(const Tensor& x, const c10::optional<Tensor>& residual, const Tensor& peer_bases, int64_t mc_base,
                             int64_t local_base, int64_t flag_off, int64_t data_off, int64_t half_bytes, Tensor state,
                             int64_t rank, int64_t world) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && (x.scalar_type() == at::kBFloat16 || x.scalar_type() == at::kFloat));
  TORCH_CHECK(state.scalar_type() == at::kInt && state.numel() >= 2 + 1024);
  if (residual) TORCH_CHECK(residual->is_contiguous() && residual->scalar_type() == x.scalar_type() && residual->numel() == x.numel());
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  const long nbytes = x.numel() * x.element_size();
  int ctas = (int)std::max<long>(1, std::min<long>(64, nbytes / 16384));
  nxd::nvls_allreduce(x.data_ptr(), residual ? residual->data_ptr() : nullptr, out.data_ptr(), peer_bases.data_ptr<int64_t>(),
                      mc_base, local_base, flag_off, data_off, half_bytes, (uint32_t*)state.data_ptr(), (int)rank, (int)world,
                      x.numel(), dt_code(x), ctas, stream());
  return out;
}
// y[M<=8, N] = sum over ranks of x[M, K] · w[N, K]ᵀ (+ residual): GEMV fused with the in-switch all-reduce
static Tensor gemv_allreduce(const Tensor& x, const Tensor& w, const c10::optional<Tensor>& residual, const Tensor& peer_bases,
                             int64_t mc_base, int64_t local_base, int64_t flag_off, int64_t data_off, int64_t half_bytes, Tensor state,
                             int64_t rank, int64_t world) {
  CHECK_IN(x); CHECK_IN(w);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1));
  TORCH_CHECK(state.scalar_type() == at::kInt && state.numel() >= 2 + 1024);
  const int M = (int)x.size(0), N = (int)w.size(0), K = (int)x.size(1);
  if (residual) TORCH_CHECK(residual->is_contiguous() && residual->scalar_type() == at::kBFloat16 && residual->numel() == (int64_t)M * N);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({M, N}, x.options());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  nxd::gemv_allreduce(x.data_ptr(), w.data_ptr(), residual ? residual->data_ptr() : nullptr, y.data_ptr(), M, N, K,
                      peer_bases.data_ptr<int64_t>(), mc_base, local_base, flag_off, data_off, half_bytes, (uint32_t*)state.data_ptr(),
                      (int)rank, (int)world, sms * 4, stream());
  return y;
}
static void nvls_all_gather(const Tensor& x, const c10::optional<Tensor>& out, const Tensor& peer_bases, int64_t mc_base,
                            int64_t local_base, int64_t flag_off, int64_t data_off, int64_t half_bytes, Tensor state, int64_t rank,
                            int64_t world, int64_t ctas) {
  CHECK_IN(x);
  c10::cuda::CUDAGuard guard(x.device());
  const long bytes = x.numel() * x.element_size();
  if (out) TORCH_CHECK(out->is_contiguous() && (long)(out->numel() * out->element_size()) == bytes * world);
  nxd::nvls_all_gather(x.data_ptr(), out ? out->data_ptr() : nullptr, peer_bases.data_ptr<int64_t>(), mc_base, local_base, flag_off,
                       data_off, half_bytes, (uint32_t*)state.data_ptr(), (int)rank, (int)world, bytes, (int)ctas, stream());
}
static void nvls_publish(const Tensor& x, const Tensor& peer_bases, int64_t mc_base, int64_t local_base, int64_t flag_off,
                         int64_t data_off, int64_t half_bytes, Tensor state, int64_t rank, int64_t world, int64_t parity, int64_t ctas) {
  CHECK_IN(x);
  c10::cuda::CUDAGuard guard(x.device());
  nxd::nvls_publish(x.data_ptr(), peer_bases.data_ptr<int64_t>(), mc_base, local_base, flag_off, data_off, half_bytes,
                    (uint32_t*)state.data_ptr(), (int)rank, (int)world, x.numel() * x.element_size(), (int)parity, (int)ctas, stream());
}
// A tensor over memory this process has mapped (a peer's symmetric slot): the caller guarantees the mapping outlives the view.
static Tensor ptr_view(int64_t ptr, const std::vector<int64_t>& shape, py::object dtype) {
  auto st = torch::python::detail::py_object_to_dtype(dtype);
  int dev; cudaGetDevice(&dev);
  return at::from_blob(reinterpret_cast<void*>(ptr), shape, at::TensorOptions().dtype(st).device(at::kCUDA, dev));
}
// x = [world, chunk…] contiguous: chunk p goes to rank p; returns [world, chunk…] with chunk p received from rank p
static Tensor nvls_all_to_all(const Tensor& x, const Tensor& peer_bases, int64_t mc_base, int64_t local_base, int64_t flag_off,
                              int64_t data_off, int64_t half_bytes, Tensor state, int64_t rank, int64_t world, int64_t ctas) {
  CHECK_IN(x);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  nxd::nvls_all_to_all(x.data_ptr(), out.data_ptr(), peer_bases.data_ptr<int64_t>(), mc_base, local_base, flag_off, data_off,
                       half_bytes, (uint32_t*)state.data_ptr(), (int)rank, (int)world, x.numel() * x.element_size(), (int)ctas, stream());
  return out;
}
// table [rows_per_rank, H] (this rank's vocab shard), ids [ntok] int64 GLOBAL vocabulary ids of this rank's tokens → [ntok, H]
static Tensor nvls_embedding_gather(const Tensor& table, const Tensor& ids, const Tensor& peer_bases, int64_t mc_base,
                                    int64_t local_base, int64_t flag_off, int64_t data_off, int64_t half_bytes, Tensor state,
                                    int64_t rank, int64_t world, int64_t ctas) {
  CHECK_IN(table);
  TORCH_CHECK(table.dim() == 2 && ids.is_cuda() && ids.scalar_type() == at::kLong && ids.is_contiguous());
  c10::cuda::CUDAGuard guard(table.device());
  const long row_bytes = table.size(1) * table.element_size();
  Tensor out = at::empty({ids.numel(), table.size(1)}, table.options());
  nxd::nvls_embedding_gather(table.data_ptr(), ids.data_ptr<long>(), out.data_ptr(), peer_bases.data_ptr<int64_t>(), mc_base,
                             local_base, flag_off, data_off, half_bytes, (uint32_t*)state.data_ptr(), (int)rank, (int)world,
                             table.size(0), row_bytes, ids.numel(), (int)ctas, stream());
  return out;
}
static Tensor nvls_reduce_scatter(const Tensor& x, const Tensor& peer_bases, int64_t mc_base, int64_t local_base, int64_t flag_off,
                                  int64_t data_off, int64_t half_bytes, Tensor state, int64_t rank, int64_t world, int64_t ctas) {
  CHECK_IN(x);
  TORCH_CHECK(x.numel() % world == 0);
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty({x.numel() / world}, x.options());
  nxd::nvls_reduce_scatter(x.data_ptr(), out.data_ptr(), peer_bases.data_ptr<int64_t>(), mc_base, local_base, flag_off, data_off,
                           half_bytes, (uint32_t*)state.data_ptr(), (int)rank, (int)world, x.numel() / world, dt_code(x), (int)ctas,
                           stream());
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("add_rmsnorm_fwd", &add_rmsnorm_fwd);
  m.def("add_rmsnorm_bwd", &add_rmsnorm_bwd);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("rope_apply", &rope_apply);
  m.def("ce_stats", &ce_stats);
  m.def("ce_backward", &ce_backward);
  m.def("multi_tensor_sq_norm", &multi_tensor_sq_norm);
  m.def("multi_tensor_scale", &multi_tensor_scale);
  m.def("fused_adamw", &fused_adamw);
  m.def("gemm_bf16", &gemm_bf16);
  m.def("gemm_bf16_2cta", &gemm_bf16_2cta);
  m.def("oneshot_allreduce", &oneshot_allreduce);
  m.def("decode_attention", &decode_attention);
  m.def("decode_attention_partial", &decode_attention_partial);
  m.def("nvls_embedding_gather", &nvls_embedding_gather);
  m.def("nvls_publish", &nvls_publish);
  m.def("ptr_view", &ptr_view);
  m.def("nvls_all_to_all", &nvls_all_to_all);
  m.def("gemv_mx", &gemv_mx);
  m.def("gemv_mx_grouped", &gemv_mx_grouped);
  m.def("gemm_mxfp8", &gemm_mxfp8);
  m.def("gemm_f4", &gemm_f4);
  m.def("moe_block_tkg", &moe_block_tkg);
  m.def("moe_block_tkg_supported", [](int64_t T, int64_t H, int64_t E, int64_t I, int64_t K) {
    return nxd::moe_block_tkg_supported((int)T, (int)H, (int)E, (int)I, (int)K);
  });
  m.def("decode_rope_kv", &decode_rope_kv);
  m.def("moe_block_metadata", &moe_block_metadata);
  m.def("row_argmax", &row_argmax);
  m.def("row_topk", &row_topk);
  m.def("row_topk_supported", &row_topk_supported);
  m.def("gemv", &gemv, py::arg("x"), py::arg("w"), py::arg("residual") = py::none());
  m.def("gemm_fp8", &gemm_fp8);
  m.def("grouped_gemm", &grouped_gemm);
  m.def("grouped_wgrad", &grouped_wgrad);
  m.def("flash_attn_fwd", &flash_attn_fwd);
  m.def("flash_attn_bwd", &flash_attn_bwd, py::arg("go"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("o"), py::arg("lse"),
        py::arg("causal"), py::arg("scale"), py::arg("sbhd_out"), py::arg("dlse") = py::none());
  m.def("ag_gemm_bf16", &ag_gemm_bf16);
  m.def("gemm_rs_bf16", &gemm_rs_bf16);
  m.def("tp_gemm_2cta", &tp_gemm_2cta);
  m.def("zero1_reduce_scatter", &zero1_reduce_scatter);
  m.def("zero1_all_gather", &zero1_all_gather);
  m.def("symm_alloc", &symm_alloc);
  m.def("symm_open", &symm_open);
  m.def("symm_free", &nxd::symm_free);
  m.def("symm_view", &symm_view);
  m.def("vmm_begin", &vmm_begin);
  m.def("vmm_send", &nxd::vmm_send);
  m.def("vmm_recv", &nxd::vmm_recv);
  m.def("vmm_bind", &nxd::vmm_bind);
  m.def("vmm_ptrs", &vmm_ptrs);
  m.def("vmm_free", &nxd::vmm_free);
  m.def("vmm_view", &vmm_view);
  m.def("tp_gemm_nvls", &tp_gemm_nvls);
  m.def("nvls_allreduce", &nvls_allreduce, py::arg("x"), py::arg("residual"), py::arg("peer_bases"), py::arg("mc_base"),
        py::arg("local_base"), py::arg("flag_off"), py::arg("data_off"), py::arg("half_bytes"), py::arg("state"), py::arg("rank"),
        py::arg("world"));
  m.def("gemv_allreduce", &gemv_allreduce, py::arg("x"), py::arg("w"), py::arg("residual"), py::arg("peer_bases"), py::arg("mc_base"),
        py::arg("local_base"), py::arg("flag_off"), py::arg("data_off"), py::arg("half_bytes"), py::arg("state"), py::arg("rank"),
        py::arg("world"));
  m.def("nvls_all_gather", &nvls_all_gather);
  m.def("nvls_reduce_scatter", &nvls_reduce_scatter);
}

#include <cstdio>
#include <cstdint>
#include <cute/arch/mma_sm100_desc.hpp>
constexpr int BM = 128, BN = 128;
constexpr uint32_t make_idesc_mx(int a_fmt, int b_fmt) {
  return ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24);
}
uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_smem_desc_noswizzle(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}
// csrc/gemm_mxf4_sm100.cu
template <int VS>
constexpr uint32_t make_idesc_f4() {
  return (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((VS == 32 ? 1u : 0u) << 23) | ((uint32_t)(BM >> 4) << 24);
}
int main() {
  int bad = 0;
  {
    using namespace cute;
    // the factory CUTLASS uses for its own kind::mxf4 / kind::mxf4nvf4 atoms (packed float_e2m1_t operands, K-major)
    auto mx = UMMA::make_instr_desc_block_scaled<float_e2m1_t, float_e2m1_t, float, float_ue8m0_t, BM, BN, UMMA::Major::K, UMMA::Major::K>();
    auto nv = UMMA::make_instr_desc_block_scaled<float_e2m1_t, float_e2m1_t, float, float_ue4m3_t, BM, BN, UMMA::Major::K, UMMA::Major::K>();
    for (uint32_t id : {0u, 2u}) {
      mx.a_sf_id_ = id; mx.b_sf_id_ = id;
      const uint32_t mine = make_idesc_f4<32>() | (id << 4) | (id << 29);
      if ((uint32_t)mx != mine) { printf("mxf4 idesc mismatch id=%u: %08x vs %08x\n", id, (uint32_t)mx, mine); ++bad; }
    }
    if ((uint32_t)nv != make_idesc_f4<16>()) { printf("nvf4 idesc mismatch: %08x vs %08x\n", (uint32_t)nv, make_idesc_f4<16>()); ++bad; }
    printf("kind::mxf4 idesc %08x, kind::mxf4nvf4 idesc %08x (MXF4Format::E2M1 = %d, UE4M3 scale = %d)\n", make_idesc_f4<32>(),
           make_idesc_f4<16>(), (int)UMMA::MXF4Format::E2M1, (int)UMMA::ScaleFormat::UE4M3);
  }
  for (int af : {0, 1, 5}) for (int bf : {0, 1, 5}) for (int k = 0; k < 4; ++k) {
    cute::UMMA::InstrDescriptorBlockScaled d{};
    d.desc_ = 0;
    d.a_format_ = af; d.b_format_ = bf; d.scale_format_ = 1; d.m_dim_ = BM >> 4; d.n_dim_ = BN >> 3;
    d.a_major_ = 0; d.b_major_ = 0; d.a_sf_id_ = k; d.b_sf_id_ = k; d.k_size_ = 0;
    const uint32_t mine = make_idesc_mx(af, bf) | ((uint32_t)k << 4) | ((uint32_t)k << 29);
    if ((uint32_t)d != mine) { printf("idesc mismatch af=%d bf=%d k=%d: %08x vs %08x\n", af, bf, k, (uint32_t)d, mine); ++bad; }
  }
  for (uint32_t addr : {0x400u, 0x30200u, 0x3fff0u}) {
    cute::UMMA::SmemDescriptor s{};
    s.desc_ = 0;
    s.start_address_ = addr >> 4; s.leading_byte_offset_ = 16 >> 4; s.stride_byte_offset_ = 128 >> 4; s.version_ = 1;
    s.layout_type_ = (uint8_t)cute::UMMA::LayoutType::SWIZZLE_NONE;
    if (s.desc_ != make_smem_desc_noswizzle(addr, 16, 128)) { printf("smem desc (none) mismatch %llx vs %llx\n", (unsigned long long)s.desc_, (unsigned long long)make_smem_desc_noswizzle(addr, 16, 128)); ++bad; }
    cute::UMMA::SmemDescriptor t{};
    t.desc_ = 0;
    t.start_address_ = addr >> 4; t.leading_byte_offset_ = 16 >> 4; t.stride_byte_offset_ = 1024 >> 4; t.version_ = 1;
    t.layout_type_ = (uint8_t)cute::UMMA::LayoutType::SWIZZLE_128B;
    if (t.desc_ != make_smem_desc_sw128(addr, 16, 1024)) { printf("smem desc (128B) mismatch %llx vs %llx\n", (unsigned long long)t.desc_, (unsigned long long)make_smem_desc_sw128(addr, 16, 1024)); ++bad; }
  }
  printf("E2M1 format code = %d, E4M3 = %d, E5M2 = %d, E8M0 scale = %d\n", (int)cute::UMMA::MXF8F6F4Format::E2M1, (int)cute::UMMA::MXF8F6F4Format::E4M3, (int)cute::UMMA::MXF8F6F4Format::E5M2, (int)cute::UMMA::ScaleFormat::UE8M0);
  printf(bad ? "MISMATCHES: %d\n" : "all descriptors agree (%d)\n", bad);
  return bad;
}

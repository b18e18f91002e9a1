// Test infrastructure, never shipped: the rest of the C ABI of include/b2t.h for the combined emulated library
// (oracle/_cache/libb2t_emu.so = this file + trace_emu.cpp + preamble_emu.cpp + field_emu.cpp, all entry points on HOST
// arrays).  tests/test_product_on_emulated_library_cpu.py points kimimaro_b200 at it (B2T_LIB-style) with CPU tensors,
// so that the product's whole Python host path runs in the CPU suite against the oracle.
//   library-level calls (capi.cu)        trivial stand-ins
//   K1 (edt.cu)                          its ring accessors are inline PTX, so the entry points run the SAME column-pass
//                                        templates (edt_fh3.cuh) through the host context of fh3_host.cpp: the hybrid
//                                        (stencil_column_v2 + column_range) for uint32 labels with integer anisotropy like
//                                        b2t_edt_ws, the envelope pass otherwise like b2t_edt
#define B2T_HOST_EMU 1
#include "emu_include/simt_impl.h"
#include "../../include/b2t.h"
#include "fh3_host.cpp"

#define EMU_EXPORT extern "C" __attribute__((visibility("default")))

EMU_EXPORT int b2t_version(void) { return 100; }
EMU_EXPORT const char* b2t_last_error(void) { return emu_last_error(); }
EMU_EXPORT int b2t_device_check(void) { return 0; }
EMU_EXPORT unsigned long long b2t_launch_count(int) { return 1; }
EMU_EXPORT int b2t_set_launch_limits(int, int) { return 0; }
EMU_EXPORT int b2t_edt_config(int, int, int, int, int) { return 0; }
EMU_EXPORT int b2t_edt_config_hybrid(int, int, int, int, int, int) { return 0; }
EMU_EXPORT int b2t_edt_config_roles(int, int, float) { return 0; }
EMU_EXPORT int b2t_edt_config_envelope(int, int) { return 0; }
EMU_EXPORT long long b2t_edt_config_xpass(int) { return 0; }
EMU_EXPORT size_t b2t_edt_workspace_bytes(int64_t sx, int64_t sy, int64_t sz) {
  return (sx <= 0 || sy <= 0 || sz <= 0) ? 0 : (size_t)sx * sy * sz * sizeof(float) + 64;
}

namespace {
template <typename T>
void edt_envelope(const T* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int bb, int ndim, float* out) {
  pass_x<T>(labels, out, sx, sy * sz, wx, bb);
  pass<T, 16, 32, 4>(labels, out, (int)sy, sx, sx, 1, sz, sx * sy, wy, bb, ndim == 2, nullptr);
  if (ndim == 3) pass<T, 16, 32, 4>(labels, out, (int)sz, sx * sy, sx, 1, sy, sx, wz, bb, 1, nullptr);
}
bool small_int(float w) { return w >= 1.0f && w <= 2048.0f && w == rintf(w); }
}  // namespace

EMU_EXPORT int b2t_edt(const void* labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                       int black_border, int ndim, float* out, void*) {
  if (sy > fh3::kMaxN || sz > fh3::kMaxN) return -1;
  const int bb = black_border ? 1 : 0;
  switch (label_bytes) {
    case 1: edt_envelope<uint8_t>((const uint8_t*)labels, sx, sy, sz, wx, wy, wz, bb, ndim, out); break;
    case 2: edt_envelope<uint16_t>((const uint16_t*)labels, sx, sy, sz, wx, wy, wz, bb, ndim, out); break;
    case 4: edt_envelope<uint32_t>((const uint32_t*)labels, sx, sy, sz, wx, wy, wz, bb, ndim, out); break;
    case 8: edt_envelope<unsigned long long>((const unsigned long long*)labels, sx, sy, sz, wx, wy, wz, bb, ndim, out); break;
    default: return -1;
  }
  return 0;
}

EMU_EXPORT int b2t_edt_ws(const void* labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                          int black_border, int ndim, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (label_bytes == 4 && ws && small_int(wx) && small_int(wy) && small_int(wz) && sy <= fh3::kMaxN && sz <= fh3::kMaxN) {
    g_stencil_v2 = 1;                         // the shipped default
    run_hybrid<10, 4, 4, 11>((const uint32_t*)labels, sx, sy, sz, wx, wy, wz, black_border ? 1 : 0, ndim, out, nullptr);
    return 0;
  }
  (void)ws_bytes;
  return b2t_edt(labels, label_bytes, sx, sy, sz, wx, wy, wz, black_border, ndim, out, stream);
}

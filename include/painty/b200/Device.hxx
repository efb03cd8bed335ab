/**
 * painty_b200 — C++ façade support: process-wide device context, status -> exception translation and the
 * host-mirror bookkeeping shared by the drop-in renderer headers in include/painty/renderer/.
 *
 * The façade keeps painty's boundary types (painty::vec, painty::Mat from the reference's own
 * painty/core/Vec.hxx and painty/image/Mat.hxx) and forwards to the C ABI in painty_b200.h.
 * Environment: PAINTY_B200_DEVICE (default 0), PAINTY_B200_PRECISION = f32 (default) | f64.
 */
#pragma once
#include <cstdlib>
#include <cstring>
#include <ios>
#include <memory>
#include <stdexcept>
#include <string>

#include "painty_b200.h"

namespace painty {
namespace b200 {

// The reference reports errors with C++ exceptions: invalid_argument from the KM / mixer
// argument checks, runtime_error elsewhere. The C ABI prefixes messages of the former with "invalid_argument".
inline void check(int rc) {
  if (rc == 0) return;
  const std::string msg = pb_last_error();
  if (msg.rfind("invalid_argument", 0) == 0) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

inline pb_context* context() {
  struct Holder {
    pb_context* ctx = nullptr;
    Holder() {
      const char* dev  = std::getenv("PAINTY_B200_DEVICE");
      const char* prec = std::getenv("PAINTY_B200_PRECISION");
      check(pb_context_create(dev ? std::atoi(dev) : 0, (prec && std::strcmp(prec, "f64") == 0) ? PB_F64 : PB_F32, &ctx));
    }
    ~Holder() { pb_context_destroy(ctx); }
  };
  static Holder h;
  return h.ctx;
}

}  // namespace b200
}  // namespace painty

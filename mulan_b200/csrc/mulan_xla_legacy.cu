// XLA GPU custom-call targets (legacy ABI with status out-parameter) over the C ABI: see
// include/mulan_b200_xla.h.  Pure host code: unpack `buffers` / `opaque`, call the entry point,
// report failure the way XLA expects.  This is the binding surface of the jaxlib the reference
// pins (jax <= 0.4.23, README.md:26), where VDM.__call__ (ldm/model_mulan_epsilon.py:280-363,
// ldm/model_mulan_velocity.py:188-268) would emit these as stablehlo.custom_call ops.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mulan_b200_xla.h"
#include "mulan_kernels.h"

namespace {

using SetFailureFn = void (*)(void* status, const char* message, size_t message_len);

// XlaCustomCallStatusSetFailure lives in the hosting process (jaxlib's xla_extension); this
// library must not link against XLA, so look it up once per call site failure (rare path).
void report(void* status, const char* where) {
  const char* msg = mulan_last_error();
  char buf[640];
  snprintf(buf, sizeof(buf), "%s: %s", where, msg);
  SetFailureFn set = nullptr;
  if (status != nullptr)
    set = reinterpret_cast<SetFailureFn>(dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure"));
  if (set != nullptr) set(status, buf, strlen(buf));
  else fprintf(stderr, "[libmulan_b200] %s\n", buf);
}

struct Unpacked {
  mulan_desc desc;
  uint32_t absent;
};

bool unpack(const char* where, const char* opaque, size_t len, Unpacked* u, void* status) {
  if (opaque == nullptr || (len != sizeof(mulan_xla_opaque) && len != sizeof(mulan_desc))) {
    char m[160];
    snprintf(m, sizeof(m), "opaque must be a mulan_xla_opaque (%zu bytes) or a mulan_desc (%zu), "
             "got %zu", sizeof(mulan_xla_opaque), sizeof(mulan_desc), len);
    mulan::set_last_error(m);
    report(status, where);
    return false;
  }
  memcpy(&u->desc, opaque, sizeof(mulan_desc));   // XLA gives no alignment guarantee
  u->absent = 0;
  if (len == sizeof(mulan_xla_opaque))
    memcpy(&u->absent, opaque + offsetof(mulan_xla_opaque, absent_mask), sizeof(uint32_t));
  return true;
}

template <typename T>
T* buf(void** buffers, uint32_t absent, int i) {
  return (absent >> i) & 1u ? nullptr : static_cast<T*>(buffers[i]);
}

bool unpack_aux(const char* where, const char* opaque, size_t len, mulan_xla_aux_opaque* o,
                void* status) {
  if (opaque == nullptr || len != sizeof(mulan_xla_aux_opaque)) {
    char m[128];
    snprintf(m, sizeof(m), "opaque must be a mulan_xla_aux_opaque (%zu bytes), got %zu",
             sizeof(mulan_xla_aux_opaque), len);
    mulan::set_last_error(m);
    report(status, where);
    return false;
  }
  memcpy(o, opaque, sizeof(*o));
  return true;
}

}  // namespace

#define F(i) buf<float>(buffers, u.absent, i)
#define CF(i) buf<const float>(buffers, u.absent, i)
#define CX(i) buf<const uint8_t>(buffers, u.absent, i)

extern "C" {

void mulan_xla_fwd_pre(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                       void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_fwd_pre", opaque, opaque_len, &u, status)) return;
  if (mulan_fwd_pre(&u.desc, CX(0), CF(1), CF(2), CF(3), CF(4), CF(5), CF(6), F(7), F(8), F(9),
                    F(10), F(11), F(12), stream))
    report(status, "mulan_xla_fwd_pre");
}

void mulan_xla_fwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                        void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_fwd_post", opaque, opaque_len, &u, status)) return;
  if (mulan_fwd_post(&u.desc, CX(0), CF(1), CF(2), CF(3), CF(4), CF(5), CF(6), CF(7), F(8),
                     stream))
    report(status, "mulan_xla_fwd_post");
}

void mulan_xla_bwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                        void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_bwd_post", opaque, opaque_len, &u, status)) return;
  if (mulan_bwd_post(&u.desc, CX(0), CF(1), CF(2), CF(3), CF(4), CF(5), CF(6), CF(7), CF(8), F(9),
                     stream))
    report(status, "mulan_xla_bwd_post");
}

void mulan_xla_fwd_bwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_fwd_bwd_post", opaque, opaque_len, &u, status)) return;
  if (mulan_fwd_bwd_post(&u.desc, CX(0), CF(1), CF(2), CF(3), CF(4), CF(5), CF(6), CF(7), CF(8),
                         F(9), F(10), stream))
    report(status, "mulan_xla_fwd_bwd_post");
}

void mulan_xla_bwd_pre(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                       void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_bwd_pre", opaque, opaque_len, &u, status)) return;
  if (mulan_bwd_pre(&u.desc, CX(0), CF(1), CF(2), CF(3), CF(4), CF(5), CF(6), CF(7), CF(8), CF(9),
                    F(10), F(11), F(12), stream))
    report(status, "mulan_xla_bwd_pre");
}

void mulan_xla_bpd_reduce(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                          void* status) {
  Unpacked u;
  if (!unpack("mulan_xla_bpd_reduce", opaque, opaque_len, &u, status)) return;
  if (mulan_bpd_reduce(&u.desc, CF(0), CF(1), CF(2), CF(3), CF(4), F(5), F(6), nullptr, stream))
    report(status, "mulan_xla_bpd_reduce");
}

#undef F
#undef CF
#undef CX

void mulan_xla_aux_topk_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status) {
  mulan_xla_aux_opaque o;
  if (!unpack_aux("mulan_xla_aux_topk_fwd", opaque, opaque_len, &o, status)) return;
  if (mulan_aux_topk_fwd(o.rows, o.latent, o.k, buf<const float>(buffers, o.absent_mask, 0),
                         buf<const float>(buffers, o.absent_mask, 1),
                         buf<float>(buffers, o.absent_mask, 2),
                         buf<float>(buffers, o.absent_mask, 3), stream))
    report(status, "mulan_xla_aux_topk_fwd");
}

void mulan_xla_aux_topk_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status) {
  mulan_xla_aux_opaque o;
  if (!unpack_aux("mulan_xla_aux_topk_bwd", opaque, opaque_len, &o, status)) return;
  if (mulan_aux_topk_bwd(o.rows, o.latent, o.k, buf<const float>(buffers, o.absent_mask, 0),
                         buf<const float>(buffers, o.absent_mask, 1),
                         buf<const float>(buffers, o.absent_mask, 2),
                         buf<const float>(buffers, o.absent_mask, 3),
                         buf<float>(buffers, o.absent_mask, 4), stream))
    report(status, "mulan_xla_aux_topk_bwd");
}

}  // extern "C"

/*
 * mulan_b200_xla.h -- XLA GPU custom-call targets over the C ABI of mulan_b200.h.
 *
 * The reference (s-sahoo/MuLAN) pins jax <= 0.4.23 (README.md:26, requirements.txt:1).  That
 * jaxlib has no typed FFI; what it CAN bind is the legacy GPU custom-call ABI with a status
 * out-parameter (xla::CustomCallApiVersion::API_VERSION_STATUS_RETURNING):
 *
 *     void target(CUstream stream, void** buffers, const char* opaque, size_t opaque_len,
 *                 XlaCustomCallStatus* status);
 *
 * registered from Python with
 *     xla_client.register_custom_call_target(b"mulan_xla_fwd_pre",
 *         PyCapsule(addr, b"xla._CUSTOM_CALL_TARGET"), platform="CUDA")
 * and emitted with jaxlib.hlo_helpers.custom_call(..., backend_config=opaque,
 * api_version=2) from an mlir lowering rule (jax_binding/mulan_jax_legacy.py shows the
 * primitive + jax.custom_vjp).  These targets live in
 * libmulan_b200.so itself, so unlike the typed-FFI handlers of jax_binding/mulan_xla_ffi.cc
 * they are compiled and tested here (through ctypes, with torch-owned device buffers standing
 * in for XLA's).
 *
 * Conventions (all targets):
 *   buffers  operand device pointers in call order, then result device pointers (XLA flattens
 *            a tuple result the same way).  XLA never passes NULL, so an operand the C ABI
 *            treats as optional is marked absent in the opaque's mask instead.
 *   opaque   the bytes of one mulan_xla_opaque (little-endian, as the struct lies in memory);
 *            `desc.rows` is filled by the binding from the operand shape.  opaque_len must be
 *            sizeof(mulan_xla_opaque) (56) or sizeof(mulan_desc) (48: no operand absent).
 *   status   on failure the message of mulan_last_error() is handed to
 *            XlaCustomCallStatusSetFailure(status, msg, len), looked up at run time in the
 *            hosting process (jaxlib's xla_extension provides it); without that symbol, or
 *            with status == NULL, the message goes to stderr and stays in mulan_last_error().
 *            Success leaves `status` untouched (XLA treats that as OK).
 *   threading / ownership: exactly the C ABI's -- enqueue-only on `stream`, no allocation, no
 *            global state; XLA calls from the executor thread of the device owning the buffers.
 */
#ifndef MULAN_B200_XLA_H_
#define MULAN_B200_XLA_H_

#include <stddef.h>
#include <stdint.h>

#include "mulan_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mulan_xla_opaque {
  mulan_desc desc;        /* 48 bytes (ABI v2: + flags, noise_rows) */
  uint32_t absent_mask;   /* bit i set: buffers[i] is to be read as NULL (optional operand) */
  uint32_t reserved;      /* 0 */
} mulan_xla_opaque;

/* Opaque of the auxiliary-latent targets (no mulan_desc there). */
typedef struct mulan_xla_aux_opaque {
  int32_t rows, latent, k;
  uint32_t absent_mask;
} mulan_xla_aux_opaque;

/* mulan_fwd_pre.   buffers: x, a, b, c, t, eps0, eps | z_t, g_net, w_save, loss_recon,
 *                           loss_klz_prior, var_sums            (w_save may be marked absent) */
void mulan_xla_fwd_pre(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                       void* status);
/* mulan_fwd_post.  buffers: x, a, b, c, t, eps, net, w_save | loss_diff */
void mulan_xla_fwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                        void* status);
/* mulan_bwd_post.  buffers: x, a, b, c, t, eps, net, w_save, gL | n_bar */
void mulan_xla_bwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                        void* status);
/* mulan_fwd_bwd_post.  buffers: x, a, b, c, t, eps, net, w_save, gL | loss_diff, n_bar */
void mulan_xla_fwd_bwd_post(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status);
/* mulan_bwd_pre.   buffers: x, a, b, c, t, eps, net, z_bar, g_bar, gL | a_bar, b_bar, c_bar
 *                  (net, z_bar, g_bar, gL may be marked absent) */
void mulan_xla_bwd_pre(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                       void* status);
/* mulan_bpd_reduce (reduce_ws == NULL: XLA result buffers are not zero-initialised, so the
 *                   single-CTA form runs; same bits as the parallel one).
 *                   buffers: loss_recon, loss_klz_prior, kl_z, loss_diff, var_sums |
 *                             scalars[6], loss_klz_total[B]      (kl_z may be marked absent) */
void mulan_xla_bpd_reduce(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                          void* status);
/* mulan_aux_topk_fwd (opaque: mulan_xla_aux_opaque).
 *                  buffers: logits, gamma_draw | embedding, kl_z   (gamma_draw may be absent) */
void mulan_xla_aux_topk_fwd(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status);
/* mulan_aux_topk_bwd (opaque: mulan_xla_aux_opaque).
 *                  buffers: logits, gamma_draw, emb_bar, klz_bar | logits_bar */
void mulan_xla_aux_topk_bwd(void* stream, void** buffers, const char* opaque, size_t opaque_len,
                            void* status);

#ifdef __cplusplus
}
#endif
#endif  /* MULAN_B200_XLA_H_ */

// cpfft_b200: fast path of the spectral operator G_K_dF (G_K_dF.f:11-87) for power-of-two grids
// (N = 16 .. 512) and the 5-smooth weak-scaling grids 320 and 400.  Same mathematics as spectral.cu, built
// from the register butterflies of fft_core.cuh:
//
//   k_fz  : [K4 : x contraction fused on load] two real z lines per CTA, each packed as an N/2
//           complex line (even/odd samples), N/2-point FFT, untangled to bins kz = 0..N/2-1.
//           The Nyquist bin is never stored: the even-N convention zeroes Ghat there.
//   k_fy  : y lines (forward or inverse), tile of TZ consecutive kz per CTA; first stage loads
//           from global memory, last stage stores to it, one shared-memory exchange between.
//   k_fx  : x forward -> Ghat contraction from integer frequencies -> x inverse, one tensor row
//           (3 components) x TZ kz per CTA; the spectrum stays digit-reversed in shared memory
//           between the two transforms, so nothing is reordered.
//   k_iz  : half-spectrum lines -> packed N/2 inverse transform -> two real samples per thread.
//
// Spectrum layout: spec[((c * NX + x) * NY + y) * NZ + kz], NZ = N/2, complex double.
#include "spectral_pow2_decl.cuh"

int cpf_pow2_init(cpfft_handle* h) {
  switch (h->N) {
    case 16: return init_pow2<16>(h);
    case 32: return init_pow2<32>(h);
    case 64: return init_pow2<64>(h);
    case 128: return init_pow2<128>(h);
    case 256: return init_pow2<256>(h);
    case 512: return init_pow2<512>(h);
    case 40: return init_pow2<40>(h);
    case 80: return init_pow2<80>(h);
    case 200: return init_pow2<200>(h);
    case 320: return init_pow2<320>(h);
    case 400: return init_pow2<400>(h);
    case 15: return init_pow2<15>(h);
    case 51: return init_pow2<51>(h);
    case 255: return init_pow2<255>(h);
  }
  return CPFFT_ERR_USAGE;
}

bool cpf_pow2_supported(int N) {
  return N == 16 || N == 32 || N == 64 || N == 128 || N == 256 || N == 512 || N == 320 || N == 400 || N == 40 || N == 80 || N == 200 ||
         N == 15 || N == 51 || N == 255;       // odd: the reference-faithful grids
}

template <int N> struct Tag {};
template <class F> static int dispatch_pow2(cpfft_handle* h, F f) {
  switch (h->N) {
    case 16: return f(Tag<16>());
    case 32: return f(Tag<32>());
    case 64: return f(Tag<64>());
    case 128: return f(Tag<128>());
    case 256: return f(Tag<256>());
    case 512: return f(Tag<512>());
    case 40: return f(Tag<40>());
    case 80: return f(Tag<80>());
    case 200: return f(Tag<200>());
    case 320: return f(Tag<320>());
    case 400: return f(Tag<400>());
    case 15: return f(Tag<15>());
    case 51: return f(Tag<51>());
    case 255: return f(Tag<255>());
  }
  cpf_set_error(h, "power-of-two spectral path called with an unsupported N");
  return CPFFT_ERR_USAGE;
}
template <int N> static int call_apply(Tag<N>, cpfft_handle* h, double* src, double* dst, bool flgK, double sc, const CgFuse* cg) {
  return apply_pow2<N>(h, src, dst, flgK, sc, cg);
}

int cpf_apply_G_pow2(cpfft_handle* h, const double* src, double* dst, bool flgK, double scale_out) {
  return dispatch_pow2(h, [&](auto tag) { return call_apply(tag, h, const_cast<double*>(src), dst, flgK, scale_out, nullptr); });
}

// One CG operator application q = G K4 p with fused direction update and dot product:
//   update_p: p <- r + beta p before the product;  on return d_partials[0 .. nparts) hold the
//   per-CTA partial sums of p.q (nparts returned).
//   x != nullptr (with update_p): the pending solution update x += (rr_alpha / *pq) p_old rides along.
int cpf_cg_apply_pow2(cpfft_handle* h, double* p, double* q, const double* r, double beta, bool update_p, int* nparts,
                      double* x, double rr_alpha, const double* pq) {
  CgFuse cg{r, beta, update_p, 0, x, rr_alpha, pq};
  *nparts = h->nxloc * h->N;   // upper bound; apply_pow2 reports the exact count of z-pass CTAs
  const int rc = dispatch_pow2(h, [&](auto tag) { return call_apply(tag, h, p, q, true, 1.0, &cg); });
  *nparts = cg.nparts;
  return rc;
}

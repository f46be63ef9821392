// Host-side transform-kit tables: the data tfm%init() builds
// (/root/reference/src/submodules/mlegs_spectfm_init.f90:6-154).  They are built once on the
// host and stay resident in HBM.  The reference runs the associated-Legendre three-term
// recurrence in FM 1.4 at 50 decimal digits and rounds to double at the end; here the same
// recurrence runs in IEEE binary128 (113-bit significand, exponent range 1e+-4932), which
// rounds to the same doubles and has the range the un-normalised values need.
#include <quadmath.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "mlegs_internal.h"

namespace mlegs {

static const double kPi = std::acos(-1.0);  // modules/mlegs_envir.f90:17

// sinit:181-214 -- Newton iteration in double, eps = 1e-15
static void gauss_legendre(int n, double *x, double *w) {
  const double eps = 1.0e-15;
  const double xm = 0.0, xl = 1.0;
  int m = (int)std::ceil((n + 1) / 2.0);
  for (int i = 1; i <= m; ++i) {
    double z = std::cos(kPi * (i - 0.25) / (n + 0.5));
    double z1 = z + 1.0, pp = 0.0;
    while (std::fabs(z - z1) > eps) {
      double p1 = 1.0, p2 = 0.0, p3;
      for (int j = 1; j <= n; ++j) {
        p3 = p2;
        p2 = p1;
        p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1);
      z1 = z;
      z = z1 - p1 / pp;
    }
    x[i - 1] = xm - xl * z;
    x[n - i] = xm + xl * z;
    w[i - 1] = 2 * xl / ((1 - z * z) * pp * pp);
    w[n - i] = w[i - 1];
  }
}

// sinit:218-250 (double precision, same summation order)
static void leg_lognorm(int ndim, int nm, double *lnrm /* (ndim, nm) col-major */) {
  int me = nm - 1;
  std::vector<double> wk(me + 1);
  wk[0] = std::log(0.5);
  for (int m = 1; m <= me; ++m)
    wk[m] = wk[m - 1] + std::log(2.0 * m + 1.0) - std::log(2.0 * m * ((2.0 * m - 1.0) * (2.0 * m - 1.0)));
  for (int m = 0; m < nm; ++m) {
    double *c = lnrm + (size_t)m * ndim;
    c[0] = wk[m];
    for (int nn = 2; nn <= ndim; ++nn) {
      int n = m + (nn - 1);
      c[nn - 1] = c[nn - 2] + std::log(2.0 * n + 1.0) - std::log(2.0 * n - 1.0) +
                  std::log(1.0 * (n - m)) - std::log(1.0 * (n + m));
    }
    for (int nn = 0; nn < ndim; ++nn) c[nn] = 0.5 * c[nn];
  }
}

// sinit:351-357
static double log_fact(int m) { return std::lgamma(2 * m + 1.0) - m * std::log(2.0) - std::lgamma(m + 1.0); }

// sinit:254-300 for one m: tbl(:, :, m) over the nx abscissae
static void leg_tbl_m(const double *x, int nx, int ne, int m, const double *lnrm_col, double *tbl /* (nx, ne) */) {
  std::vector<__float128> scale(ne);
  for (int nn = 0; nn < ne; ++nn)
    scale[nn] = expq((__float128)log_fact(m + nn) + (__float128)lnrm_col[nn]);
  std::vector<__float128> col(ne);
  for (int xx = 0; xx < nx; ++xx) {
    __float128 xv = (__float128)x[xx];
    __float128 s = sqrtq(1.0Q - xv * xv);
    __float128 p = 1.0Q;
    for (int k = 0; k < m; ++k) p *= s;
    if (m & 1) p = -p;
    col[0] = p;
    if (ne > 1) col[1] = xv * col[0];
    for (int nn = 3; nn <= ne; ++nn) {
      int n = m + nn - 1;
      col[nn - 1] = 1.0Q / (__float128)(n - m) *
                    (col[nn - 2] * xv - col[nn - 3] * (__float128)(n + m - 1) / (__float128)(2 * n - 1) /
                                            (__float128)(2 * n - 3));
    }
    for (int nn = 0; nn < ne; ++nn) tbl[(size_t)nn * nx + xx] = (double)(col[nn] * scale[nn]);
  }
}

int build_tfm_tables(const mlegs_params *p, double *x, double *w, double *ln, double *r, double *lognorm,
                     double *pf, double *at0, double *at1, double *ak) {
  const int nr = p->nr, nrh = nr / 2, ne = p->nrchop + 14, nm = p->npchop;
  gauss_legendre(nr, x, w);
  for (int i = 0; i < nr; ++i) {
    ln[i] = -std::log(1.0 - x[i]);
    r[i] = p->ell * std::sqrt((1.0 + x[i]) / (1.0 - x[i]));
  }
  // sinit:132-134
  for (int i = 0; i < p->nz; ++i) ak[i] = 2.0 * kPi / p->zlen * (double)(i - p->nz);
  for (int i = 0; i <= p->nz / 2 && i < p->nz; ++i) ak[i] = 2.0 * kPi / p->zlen * (double)i;
  leg_lognorm(ne, nm, lognorm);

  unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; ++t) {
    pool.emplace_back([=]() {
      for (int m = (int)t; m < nm; m += (int)nthreads)
        leg_tbl_m(x, nrh, ne, m, lognorm + (size_t)m * ne, pf + (size_t)m * nrh * ne);
    });
  }
  for (auto &th : pool) th.join();

  // sinit:147-150
  const double xm1 = -1.0 + 1.0e-15, xp1 = 1.0 - 1.0e-15;
  leg_tbl_m(&xm1, 1, p->nrchop, 0, lognorm, at0);
  leg_tbl_m(&xp1, 1, p->nrchop, 0, lognorm, at1);
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// On-disk cache of the expensive tables: x, w, lognorm, pf, at0, at1 depend on
// (nr, nrchop, npchop) only -- not on ell, zlen, nz or the physics -- so one file per triple serves every run.
// Layout: magic "MLEGSTB1" | int32 nr, nrchop, npchop, reserved | the six arrays as raw doubles | FNV-1a 64 of the
// payload.  Written to a temporary name and renamed, so a concurrent reader never sees a partial file; a file
// that fails any check is ignored and rebuilt.
// ---------------------------------------------------------------------------------------------------------------
static uint64_t fnv1a(const void *data, size_t n, uint64_t h) {   // FNV-1a over 64-bit words (n is a multiple of 8)
  const uint64_t *p = (const uint64_t *)data;
  for (size_t i = 0; i < n / 8; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

int build_tfm_tables_cached(const mlegs_params *p, const char *cache_dir, double *x, double *w, double *ln, double *r,
                            double *lognorm, double *pf, double *at0, double *at1, double *ak, int *from_cache) {
  if (from_cache) *from_cache = 0;
  if (!cache_dir || !cache_dir[0]) return build_tfm_tables(p, x, w, ln, r, lognorm, pf, at0, at1, ak);
  const int nr = p->nr, nrh = nr / 2, ne = p->nrchop + 14, nm = p->npchop;
  const std::string path = std::string(cache_dir) + "/mlegs_tables_" + std::to_string(nr) + "_" +
                           std::to_string(p->nrchop) + "_" + std::to_string(nm) + ".bin";
  struct Part {
    double *ptr;
    size_t n;
  };
  const Part parts[6] = {{x, (size_t)nr},   {w, (size_t)nr},           {lognorm, (size_t)ne * nm},
                         {pf, (size_t)nrh * ne * nm}, {at0, (size_t)p->nrchop}, {at1, (size_t)p->nrchop}};
  const char magic[8] = {'M', 'L', 'E', 'G', 'S', 'T', 'B', '1'};
  const int32_t hdr[4] = {nr, p->nrchop, nm, 0};
  if (FILE *fp = fopen(path.c_str(), "rb")) {
    char m2[8];
    int32_t h2[4];
    bool ok = fread(m2, 1, 8, fp) == 8 && std::equal(m2, m2 + 8, magic) && fread(h2, 4, 4, fp) == 4 &&
              std::equal(h2, h2 + 4, hdr);
    uint64_t h = 1469598103934665603ull, stored = 0;
    for (int q = 0; q < 6 && ok; ++q) {
      ok = fread(parts[q].ptr, sizeof(double), parts[q].n, fp) == parts[q].n;
      if (ok) h = fnv1a(parts[q].ptr, parts[q].n * sizeof(double), h);
    }
    ok = ok && fread(&stored, sizeof(stored), 1, fp) == 1 && stored == h;
    fclose(fp);
    if (ok) {
      // the cheap, physics-dependent tables (sinit:120-134)
      for (int i = 0; i < nr; ++i) {
        ln[i] = -std::log(1.0 - x[i]);
        r[i] = p->ell * std::sqrt((1.0 + x[i]) / (1.0 - x[i]));
      }
      for (int i = 0; i < p->nz; ++i) ak[i] = 2.0 * kPi / p->zlen * (double)(i - p->nz);
      for (int i = 0; i <= p->nz / 2 && i < p->nz; ++i) ak[i] = 2.0 * kPi / p->zlen * (double)i;
      if (from_cache) *from_cache = 1;
      return MLEGS_OK;
    }
  }
  MLEGS_TRY(build_tfm_tables(p, x, w, ln, r, lognorm, pf, at0, at1, ak));
  const std::string tmp = path + ".tmp." + std::to_string((long long)std::hash<std::thread::id>()(std::this_thread::get_id()));
  if (FILE *fp = fopen(tmp.c_str(), "wb")) {
    bool ok = fwrite(magic, 1, 8, fp) == 8 && fwrite(hdr, 4, 4, fp) == 4;
    uint64_t h = 1469598103934665603ull;
    for (int q = 0; q < 6 && ok; ++q) {
      ok = fwrite(parts[q].ptr, sizeof(double), parts[q].n, fp) == parts[q].n;
      h = fnv1a(parts[q].ptr, parts[q].n * sizeof(double), h);
    }
    ok = ok && fwrite(&h, sizeof(h), 1, fp) == 1;
    ok = (fclose(fp) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
  }
  return MLEGS_OK;   // an unwritable cache directory is not an error: the tables are built either way
}

}  // namespace mlegs

// Field I/O in the reference's file formats: msave_scalar / mload_scalar of
// /root/reference/src/submodules/mlegs_scalar_io.f90:6-250 and the assemble / disassemble pair of
// /root/reference/src/submodules/mlegs_scalar_dist.f90:70-368.
//
// The reference gathers the whole array onto rank 0 (`assemble`, MPI_Gatherv of derived types) and lets rank 0
// write it; the global array therefore has to fit one host, and every other rank idles.  Here the GLOBAL file is
// written cooperatively: the file layout is fixed by glb_sz alone (binary stream, or fixed-width 1PE24.15E3
// records), so rank 0 creates it with its header and trailer and every rank then pwrite()s the runs of its own slab
// at their byte offsets -- the file is bit-identical to the one a single rank would write.  Loading is the mirror
// image: every rank pread()s its own slab.  Host code only; fields cross PCIe once.
//
// Stream layouts (gfortran `access='stream'`, no record markers), mlegs_scalar_io.f90:47-52, 86-90:
//   global: int32 n1,n2,n3 | complex(p8) a(n1,n2,n3) column-major | real(p8) ln | int32 nrchop,npchop,nzchop offsets
//           | character(3) space
//   local : int32 glb_sz(3), loc_sz(3), loc_st(3) | complex(p8) e(loc_sz) | same trailer;  file name fn_<rank>
// Formatted layouts, :54-75, 92-114: `(3(1X,I10))` size lines, one line per (k, i) holding the n2 (re, im) pairs in
// 1PE24.15E3, a blank line between k planes, list-directed trailer lines.  List-directed `write(fo,*) ''` ends a
// record with one blank, which is reproduced; the reader is whitespace-tolerant either way.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

namespace mlegs {

int dist_allreduce(double *d_inout, int n);   // dist.cu (used here as the inter-rank barrier)

namespace {

const int NUMW = 24;   // formatted_num_str_len, modules/mlegs_envir.f90:21

// one number in Fortran 1PE24.15E3
void fmt_1pe(double v, char *out /* NUMW chars, no terminator */) {
  char tmp[64];
  if (std::isnan(v)) {
    snprintf(tmp, sizeof(tmp), "%*s", NUMW, "NaN");
  } else if (std::isinf(v)) {
    snprintf(tmp, sizeof(tmp), "%*s", NUMW, v > 0 ? "Infinity" : "-Infinity");
  } else {
    char m[48];
    snprintf(m, sizeof(m), "%.15E", v);          // d.dddddddddddddddE+dd[d]
    char *e = strchr(m, 'E');
    int ex = atoi(e + 1);
    *e = 0;
    char body[48];
    snprintf(body, sizeof(body), "%sE%c%03d", m, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    snprintf(tmp, sizeof(tmp), "%*s", NUMW, body);
  }
  memcpy(out, tmp, NUMW);
}

struct Layout {
  int n1, n2, n3;          // extents of the array the file describes (global array, or the local block)
  bool binary, global;
  size_t header;           // bytes before the data
  size_t line;             // formatted: bytes of one (k,i) line incl. its terminator
  size_t blank;            // formatted: bytes of the blank line between planes
  size_t data;             // bytes of the data section
};

std::string size_line(const int *v) {
  char b[64];
  snprintf(b, sizeof(b), " %10d %10d %10d", v[0], v[1], v[2]);
  return b;
}

Layout make_layout(const mlegs_field *s, bool binary, bool global) {
  Layout L;
  const int *sz = global ? s->glb_sz : s->loc_sz;
  L.n1 = sz[0];
  L.n2 = sz[1];
  L.n3 = sz[2];
  L.binary = binary;
  L.global = global;
  if (binary) {
    L.header = (global ? 3 : 9) * sizeof(int);
    L.line = L.blank = 0;
    L.data = (size_t)L.n1 * L.n2 * L.n3 * sizeof(cplx);
  } else {
    // global: "(3(1X,I10))" advance='no' then write(fo,*) '' -> 33 chars + " \n"; local: three advancing lines
    L.header = global ? (33 + 2) : 3 * (33 + 1);
    L.line = (size_t)L.n2 * 2 * NUMW + 2;      // numbers, then " \n" from write(fo,*) ''
    L.blank = 2;                               // " \n"
    L.data = (size_t)L.n3 * L.n1 * L.line + (size_t)(L.n3 > 0 ? L.n3 - 1 : 0) * L.blank;
  }
  return L;
}

std::string header_bytes(const mlegs_field *s, const Layout &L) {
  std::string h;
  if (L.binary) {
    if (L.global) {
      h.append((const char *)s->glb_sz, 3 * sizeof(int));
    } else {
      h.append((const char *)s->glb_sz, 3 * sizeof(int));
      h.append((const char *)s->loc_sz, 3 * sizeof(int));
      h.append((const char *)s->loc_st, 3 * sizeof(int));
    }
  } else if (L.global) {
    h = size_line(s->glb_sz) + " \n";
  } else {
    h = size_line(s->glb_sz) + "\n" + size_line(s->loc_sz) + "\n" + size_line(s->loc_st) + "\n";
  }
  return h;
}

std::string trailer_bytes(const mlegs_field *s, const Layout &L) {
  std::string t;
  if (L.binary) {
    t.append((const char *)&s->ln, sizeof(double));
    int off[3] = {s->nrchop_offset, s->npchop_offset, s->nzchop_offset};
    t.append((const char *)off, sizeof(off));
    t.append(s->space, 3);
  } else {
    char b[160];
    // write(fo,*) s%ln, offsets ; write(fo,*) s%space   (list-directed).  The field widths of list-directed output are
    // the compiler's choice (gfortran prints the real in a G25.17-style field and I12 integers), so this trailer is
    // what the reference's list-directed READ accepts, not a byte copy of what a given compiler writes; the data
    // records above follow the reference's explicit edit descriptors exactly.
    snprintf(b, sizeof(b), " %24.16E %11d %11d %11d\n %.3s\n", s->ln, s->nrchop_offset, s->npchop_offset,
             s->nzchop_offset, s->space);
    t = b;
  }
  return t;
}

int io_fail(const std::string &what, const char *fn) {
  return fail(MLEGS_E_ARG, what + " " + fn + (errno ? std::string(": ") + strerror(errno) : std::string()));
}

int pwrite_all(int fd, const void *buf, size_t n, size_t off, const char *fn) {
  const char *p = (const char *)buf;
  while (n) {
    ssize_t w = pwrite(fd, p, n, (off_t)off);
    if (w <= 0) return io_fail("msave_scalar: cannot write", fn);
    p += w;
    off += (size_t)w;
    n -= (size_t)w;
  }
  return MLEGS_OK;
}

int pread_all(int fd, void *buf, size_t n, size_t off, const char *fn) {
  char *p = (char *)buf;
  while (n) {
    ssize_t r = pread(fd, p, n, (off_t)off);
    if (r <= 0) return io_fail("mload_scalar: unexpected end of", fn);
    p += r;
    off += (size_t)r;
    n -= (size_t)r;
  }
  return MLEGS_OK;
}

std::string local_name(const char *fn, int rank) { return std::string(fn) + "_" + std::to_string(rank); }

}  // namespace

// Writes this rank's part of the file.  create != 0 (rank 0 of a global file; every rank of a per-rank file): the file
// is created/truncated to its final size and receives header and trailer first.
int msave_part(const mlegs_field *s, const cplx *host_e, const char *fn, int is_binary, int is_global, int rank,
               int create) {
  errno = 0;
  const Layout L = make_layout(s, is_binary != 0, is_global != 0);
  const std::string name = is_global ? std::string(fn) : local_name(fn, rank);
  const std::string head = header_bytes(s, L), tail = trailer_bytes(s, L);
  // formatted files end the last data line with the `write(fo,*) ''` that precedes the trailer: already counted in L.line
  const size_t total = head.size() + L.data + tail.size();
  int fd = open(name.c_str(), create ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
  if (fd < 0) return io_fail("msave_scalar: cannot open", name.c_str());
  int rc = MLEGS_OK;
  if (create) {
    if (ftruncate(fd, (off_t)total) != 0) rc = io_fail("msave_scalar: cannot size", name.c_str());
    if (rc == MLEGS_OK) rc = pwrite_all(fd, head.data(), head.size(), 0, name.c_str());
    if (rc == MLEGS_OK) rc = pwrite_all(fd, tail.data(), tail.size(), head.size() + L.data, name.c_str());
  }
  // position of the local block inside the array the file describes; local column j is global column j0 + j ms
  // (ms > 1: the cyclic azimuthal ownership of a multi-rank run, mlegs_internal.h)
  const int i0 = is_global ? s->loc_st[0] : 0, j0 = is_global ? s->loc_st[1] : 0, k0 = is_global ? s->loc_st[2] : 0;
  const int ms = (is_global && ctx().ready) ? field_mstride(s) : 1;
  const int l1 = s->loc_sz[0], l2 = s->loc_sz[1], l3 = s->loc_sz[2];
  if (L.binary) {
    // runs of l1 contiguous elements; whole planes / the whole block when the leading extents are complete
    const bool full1 = (l1 == L.n1) && ms == 1, full2 = full1 && (l2 == L.n2);
    for (int k = 0; k < l3 && rc == MLEGS_OK; ++k) {
      if (full2) {
        if (k > 0) break;
        rc = pwrite_all(fd, host_e, (size_t)l1 * l2 * l3 * sizeof(cplx),
                        head.size() + ((size_t)k0 * L.n2 * L.n1) * sizeof(cplx), name.c_str());
        break;
      }
      for (int j = 0; j < l2 && rc == MLEGS_OK; ++j) {
        const size_t src = ((size_t)k * l2 + j) * l1;
        const size_t dst = (((size_t)(k0 + k) * L.n2 + (j0 + j * ms)) * L.n1 + i0) * sizeof(cplx);
        if (full1) {   // columns j0..j0+l2-1 of plane k are one run
          rc = pwrite_all(fd, host_e + (size_t)k * l2 * l1, (size_t)l1 * l2 * sizeof(cplx), head.size() + dst,
                          name.c_str());
          break;
        }
        rc = pwrite_all(fd, host_e + src, (size_t)l1 * sizeof(cplx), head.size() + dst, name.c_str());
      }
    }
  } else {
    // line (k, i): n2 pairs; this rank owns pairs j0 .. j0+l2-1 of lines i0 .. i0+l1-1 of planes k0 .. k0+l3-1
    std::vector<char> seg((size_t)l2 * 2 * NUMW + 2);
    const bool ends_line = l2 > 0 && (j0 + (l2 - 1) * ms + 1 == L.n2);
    for (int k = 0; k < l3 && rc == MLEGS_OK; ++k) {
      const size_t plane_off = head.size() + (size_t)(k0 + k) * ((size_t)L.n1 * L.line + L.blank);
      for (int i = 0; i < l1 && rc == MLEGS_OK && ms > 1; ++i) {
        // strided columns: one pair at a time; the owner of the last column also ends the line
        for (int j = 0; j < l2 && rc == MLEGS_OK; ++j) {
          const cplx v = host_e[((size_t)k * l2 + j) * l1 + i];
          char *p = seg.data();
          fmt_1pe(v.x, p);
          fmt_1pe(v.y, p + NUMW);
          size_t n = 2 * NUMW;
          if (ends_line && j == l2 - 1) {
            seg[n++] = ' ';
            seg[n++] = '\n';
          }
          rc = pwrite_all(fd, seg.data(), n, plane_off + (size_t)(i0 + i) * L.line + (size_t)(j0 + j * ms) * 2 * NUMW,
                          name.c_str());
        }
      }
      for (int i = 0; i < l1 && rc == MLEGS_OK && ms == 1; ++i) {
        char *p = seg.data();
        for (int j = 0; j < l2; ++j) {
          const cplx v = host_e[((size_t)k * l2 + j) * l1 + i];
          fmt_1pe(v.x, p);
          fmt_1pe(v.y, p + NUMW);
          p += 2 * NUMW;
        }
        size_t n = (size_t)l2 * 2 * NUMW;
        if (ends_line) {
          seg[n++] = ' ';
          seg[n++] = '\n';
        }
        rc = pwrite_all(fd, seg.data(), n, plane_off + (size_t)(i0 + i) * L.line + (size_t)j0 * 2 * NUMW, name.c_str());
      }
      // blank line between planes: written by the rank that owns the last line's end
      if (rc == MLEGS_OK && ends_line && i0 + l1 == L.n1 && k0 + k + 1 < L.n3)
        rc = pwrite_all(fd, " \n", 2, plane_off + (size_t)L.n1 * L.line, name.c_str());
    }
  }
  if (close(fd) != 0 && rc == MLEGS_OK) rc = io_fail("msave_scalar: cannot close", name.c_str());
  return rc;
}

// Reads this rank's slab (and the metadata) from the file.  s supplies the expected sizes and the slab position.
int mload_part(const char *fn, mlegs_field *s, cplx *host_e, int is_binary, int is_global, int rank) {
  errno = 0;
  const Layout L = make_layout(s, is_binary != 0, is_global != 0);
  const std::string name = is_global ? std::string(fn) : local_name(fn, rank);
  const int i0 = is_global ? s->loc_st[0] : 0, j0 = is_global ? s->loc_st[1] : 0, k0 = is_global ? s->loc_st[2] : 0;
  const int ms = (is_global && ctx().ready) ? field_mstride(s) : 1;
  const int l1 = s->loc_sz[0], l2 = s->loc_sz[1], l3 = s->loc_sz[2];
  auto size_error = [&](const int *got) {
    char b[256];
    snprintf(b, sizeof(b),
             "mloadc: size inconsistency between data and array (data %d %d %d, array %d %d %d)", got[0], got[1], got[2],
             L.n1, L.n2, L.n3);
    return fail(MLEGS_E_ARG, b);
  };
  if (L.binary) {
    int fd = open(name.c_str(), O_RDONLY);
    if (fd < 0) return fail(MLEGS_E_ARG, std::string("mload_scalar: cannot open ") + name);
    int hdr[9];
    int rc = pread_all(fd, hdr, L.header, 0, name.c_str());
    if (rc == MLEGS_OK) {
      const int *got = is_global ? hdr : hdr + 3;
      if (got[0] != L.n1 || got[1] != L.n2 || got[2] != L.n3) rc = size_error(got);
    }
    for (int k = 0; k < l3 && rc == MLEGS_OK; ++k)
      for (int j = 0; j < l2 && rc == MLEGS_OK; ++j)
        rc = pread_all(fd, host_e + ((size_t)k * l2 + j) * l1, (size_t)l1 * sizeof(cplx),
                       L.header + (((size_t)(k0 + k) * L.n2 + (j0 + j * ms)) * L.n1 + i0) * sizeof(cplx), name.c_str());
    if (rc == MLEGS_OK) {
      char t[sizeof(double) + 3 * sizeof(int) + 3];
      rc = pread_all(fd, t, sizeof(t), L.header + L.data, name.c_str());
      if (rc == MLEGS_OK) {
        int off[3];
        memcpy(&s->ln, t, sizeof(double));
        memcpy(off, t + sizeof(double), sizeof(off));
        s->nrchop_offset = off[0];
        s->npchop_offset = off[1];
        s->nzchop_offset = off[2];
        memcpy(s->space, t + sizeof(double) + sizeof(off), 3);
        s->space[3] = 0;
      }
    }
    close(fd);
    return rc;
  }
  // formatted: token stream (whitespace-tolerant, so files written by the reference load as well)
  FILE *fp = fopen(name.c_str(), "r");
  if (!fp) return fail(MLEGS_E_ARG, std::string("mload_scalar: cannot open ") + name);
  int rc = MLEGS_OK;
  int hdr[9] = {0};
  const int nh = is_global ? 3 : 9;
  for (int q = 0; q < nh && rc == MLEGS_OK; ++q)
    if (fscanf(fp, "%d", &hdr[q]) != 1) rc = io_fail("mload_scalar: bad header in", name.c_str());
  if (rc == MLEGS_OK) {
    const int *got = is_global ? hdr : hdr + 3;
    if (got[0] != L.n1 || got[1] != L.n2 || got[2] != L.n3) rc = size_error(got);
  }
  char tok[128] = "0";
  for (int k = 0; k < L.n3 && rc == MLEGS_OK; ++k)
    for (int i = 0; i < L.n1 && rc == MLEGS_OK; ++i)
      for (int j = 0; j < L.n2 && rc == MLEGS_OK; ++j) {
        double re = 0.0, im = 0.0;
        if (fscanf(fp, "%127s", tok) != 1) rc = io_fail("mload_scalar: unexpected end of", name.c_str());
        if (rc == MLEGS_OK) re = strtod(tok, nullptr);
        if (rc == MLEGS_OK && fscanf(fp, "%127s", tok) != 1) rc = io_fail("mload_scalar: unexpected end of", name.c_str());
        if (rc == MLEGS_OK) im = strtod(tok, nullptr);
        const int li = i - i0, lk = k - k0;
        const int lj = (j >= j0 && (j - j0) % ms == 0) ? (j - j0) / ms : -1;
        if (rc == MLEGS_OK && li >= 0 && li < l1 && lj >= 0 && lj < l2 && lk >= 0 && lk < l3)
          host_e[((size_t)lk * l2 + lj) * l1 + li] = make_double2(re, im);
      }
  if (rc == MLEGS_OK) {
    int off[3];
    if (fscanf(fp, "%127s", tok) != 1) rc = io_fail("mload_scalar: missing trailer in", name.c_str());
    s->ln = strtod(tok, nullptr);
    for (int q = 0; q < 3 && rc == MLEGS_OK; ++q)
      if (fscanf(fp, "%d", &off[q]) != 1) rc = io_fail("mload_scalar: missing trailer in", name.c_str());
    if (rc == MLEGS_OK && fscanf(fp, "%127s", tok) == 1 && strlen(tok) >= 3) {
      s->nrchop_offset = off[0];
      s->npchop_offset = off[1];
      s->nzchop_offset = off[2];
      memcpy(s->space, tok, 3);
      s->space[3] = 0;
    } else if (rc == MLEGS_OK) {
      rc = io_fail("mload_scalar: missing trailer in", name.c_str());
    }
  }
  fclose(fp);
  return rc;
}

int dist_check_timeout();                     // dist.cu: stream synchronize + barrier failure report

// MPI_Barrier(comm_glb) of mlegs_scalar_io.f90:79 as a HOST barrier: the device all-reduce only completes once every
// rank has entered it, and the host waits for it here -- no rank touches the file before rank 0 has laid it out.
static int barrier_ranks() {
  Context &c = ctx();
  if (c.nranks <= 1) return MLEGS_OK;
  CUDA_TRY(cudaMemsetAsync(c.d_red, 0, sizeof(double), (cudaStream_t)c.stream));
  MLEGS_TRY(dist_allreduce(c.d_red, 1));
  return dist_check_timeout();
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_msave_part(const mlegs_field *meta, const void *host_e, const char *fn, int is_binary, int is_global,
                          int rank, int create) {
  return msave_part(meta, (const cplx *)host_e, fn, is_binary, is_global, rank, create);
}

int mlegs_b200_mload_part(const char *fn, mlegs_field *meta, void *host_e, int is_binary, int is_global, int rank) {
  return mload_part(fn, meta, (cplx *)host_e, is_binary, is_global, rank);
}

int mlegs_b200_msave(const mlegs_field *s, const char *fn, int is_binary, int is_global) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  const size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
  std::vector<cplx> h(n ? n : 1);
  CUDA_TRY(cudaMemcpyAsync(h.data(), s->e, n * sizeof(cplx), cudaMemcpyDeviceToHost, (cudaStream_t)c.stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)c.stream));
  if (!is_global || c.nranks == 1) return msave_part(s, h.data(), fn, is_binary, is_global, c.rank, 1);
  // global file, several ranks: rank 0 lays the file out, then everybody fills in its slab
  int rc = MLEGS_OK;
  if (c.rank == 0) rc = msave_part(s, h.data(), fn, is_binary, 1, 0, 1);
  MLEGS_TRY(barrier_ranks());
  if (c.rank != 0) rc = msave_part(s, h.data(), fn, is_binary, 1, c.rank, 0);
  MLEGS_TRY(barrier_ranks());   // call MPI_barrier(comm_glb), mlegs_scalar_io.f90:79
  return rc;
}

int mlegs_b200_mload(const char *fn, mlegs_field *s, int is_binary, int is_global) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  const size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
  std::vector<cplx> h(n ? n : 1);
  MLEGS_TRY(mload_part(fn, s, h.data(), is_binary, is_global, c.rank));
  CUDA_TRY(cudaMemcpyAsync(s->e, h.data(), n * sizeof(cplx), cudaMemcpyHostToDevice, (cudaStream_t)c.stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)c.stream));
  return MLEGS_OK;
}

}  // extern "C"

// Entry points that are declared in include/mlegs_b200.h but not implemented yet.
#include "kernels.h"
using namespace mlegs;
namespace mlegs { int build_operator_tables() { return MLEGS_OK; } }
#define NOT_YET(name) return fail(MLEGS_E_STATE, name ": not implemented yet")
extern "C" {
int mlegs_b200_chop(mlegs_field *) { NOT_YET("chop"); }
int mlegs_b200_dealias(mlegs_field *) { NOT_YET("dealias"); }
int mlegs_b200_svv_filter(mlegs_field *, double *) { NOT_YET("svv_filter"); }
int mlegs_b200_calcat0(const mlegs_field *, double *) { NOT_YET("calcat0"); }
int mlegs_b200_calcat1(const mlegs_field *, double *) { NOT_YET("calcat1"); }
int mlegs_b200_zeroat1(mlegs_field *) { NOT_YET("zeroat1"); }
int mlegs_b200_delsqp(mlegs_field *) { NOT_YET("delsqp"); }
int mlegs_b200_idelsqp(mlegs_field *) { NOT_YET("idelsqp"); }
int mlegs_b200_xxdx(mlegs_field *) { NOT_YET("xxdx"); }
int mlegs_b200_del2h(mlegs_field *) { NOT_YET("del2h"); }
int mlegs_b200_del2(mlegs_field *) { NOT_YET("del2"); }
int mlegs_b200_idel2(mlegs_field *, int, double) { NOT_YET("idel2"); }
int mlegs_b200_ihelm(mlegs_field *, double) { NOT_YET("ihelm"); }
int mlegs_b200_helmp(mlegs_field *, int, double, double) { NOT_YET("helmp"); }
int mlegs_b200_ihelmp(mlegs_field *, int, double, double) { NOT_YET("ihelmp"); }
int mlegs_b200_fefe(mlegs_field *, const mlegs_field *, double) { NOT_YET("fefe"); }
int mlegs_b200_febe(mlegs_field *, const mlegs_field *, double) { NOT_YET("febe"); }
int mlegs_b200_abcn(mlegs_field *, mlegs_field *, mlegs_field *, mlegs_field *, double) { NOT_YET("abcn"); }
int mlegs_b200_vecprod(mlegs_field *, mlegs_field *, mlegs_field *, const mlegs_field *, const mlegs_field *,
                       const mlegs_field *) { NOT_YET("vecprod"); }
int mlegs_b200_vec2tp(const mlegs_field *, const mlegs_field *, const mlegs_field *, mlegs_field *, mlegs_field *) {
  NOT_YET("vec2tp");
}
int mlegs_b200_tp2vec(const mlegs_field *, const mlegs_field *, mlegs_field *, mlegs_field *, mlegs_field *) {
  NOT_YET("tp2vec");
}
int mlegs_b200_tp2curlvec(const mlegs_field *, const mlegs_field *, mlegs_field *, mlegs_field *, mlegs_field *) {
  NOT_YET("tp2curlvec");
}
int mlegs_b200_axpby(mlegs_field *, double, const mlegs_field *, double) { NOT_YET("axpby"); }
int mlegs_b200_is_finite(const mlegs_field *, int *) { NOT_YET("is_finite"); }
}

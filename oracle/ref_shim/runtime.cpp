// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Runtime behind the shim headers in oracle/ref_shim/include plus a small C driver,
// so that the reference's OWN hot-path translation units (compiled unmodified from
// /root/reference/src by oracle/Makefile into oracle/_ref/) can run here without R
// or gdsfmt: an in-process model of the R objects they create, the GDS_* calls
// re-pointed at an in-memory uint8 genotype matrix, pthread wrappers, and ref_*()
// entry points that drive the reference exactly like R does
// (gnrSetGenoSpace -> gnrSelSNP_Base -> gnrGRM / gnrPCA / gnrIBSNum / ...,
// R/Internal.R:427-447, R/IBD.R:594).
#include <pthread.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "R_GDS_CPP.h"
#include "R_ext/Rdynload.h"
#include "dGenGWAS.h"

// ------------------------------------------------------------------ R model
static shim_sexp g_nil = {NILSXP, 0, nullptr, nullptr, nullptr};
static shim_sexp g_names_sym = {NILSXP, 0, nullptr, nullptr, nullptr};
static shim_sexp g_dim_sym = {NILSXP, 0, nullptr, nullptr, nullptr};
static bool g_quiet = true;

extern "C" {
SEXP R_NilValue = &g_nil, R_NamesSymbol = &g_names_sym, R_DimSymbol = &g_dim_sym;
double R_NaN = NAN, R_NaReal = NAN, R_PosInf = INFINITY, R_NegInf = -INFINITY;
int R_NaInt = INT32_MIN;

SEXP Rf_allocVector(int type, R_xlen_t n) {
    SEXP s = new shim_sexp{type, n, nullptr, R_NilValue, R_NilValue};
    size_t el = type == REALSXP ? 8 : (type == INTSXP || type == LGLSXP) ? 4 : sizeof(void *);
    s->data = calloc((size_t)(n > 0 ? n : 1), el);
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(int type, int nr, int nc) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
    SEXP d = Rf_allocVector(INTSXP, 2);
    ((int *)d->data)[0] = nr;
    ((int *)d->data)[1] = nc;
    s->dim = d;
    return s;
}
SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int) {}
double *REAL(SEXP s) { return (double *)s->data; }
int *INTEGER(SEXP s) { return (int *)s->data; }
int *LOGICAL(SEXP s) { return (int *)s->data; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; return v; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
const char *CHAR(SEXP s) { return (const char *)s->data; }
SEXP Rf_mkChar(const char *c) {
    SEXP s = new shim_sexp{CHARSXP, (R_xlen_t)strlen(c), strdup(c), R_NilValue, R_NilValue};
    return s;
}
SEXP Rf_mkString(const char *c) {
    SEXP s = Rf_allocVector(STRSXP, 1);
    ((SEXP *)s->data)[0] = Rf_mkChar(c);
    return s;
}
int Rf_asInteger(SEXP s) {
    if (s->len < 1) return R_NaInt;
    if (s->type == REALSXP) { double v = REAL(s)[0]; return v != v ? R_NaInt : (int)v; }
    if (s->type == INTSXP || s->type == LGLSXP) return INTEGER(s)[0];
    return R_NaInt;
}
int Rf_asLogical(SEXP s) {
    if (s->len < 1) return R_NaInt;
    if (s->type == REALSXP) { double v = REAL(s)[0]; return v != v ? R_NaInt : (v != 0); }
    if (s->type == INTSXP || s->type == LGLSXP) {
        int v = INTEGER(s)[0];
        return v == R_NaInt ? R_NaInt : (v != 0);
    }
    return R_NaInt;
}
double Rf_asReal(SEXP s) {
    if (s->len < 1) return R_NaReal;
    if (s->type == REALSXP) return REAL(s)[0];
    if (s->type == INTSXP || s->type == LGLSXP) {
        int v = INTEGER(s)[0];
        return v == R_NaInt ? R_NaReal : (double)v;
    }
    return R_NaReal;
}
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_duplicate(SEXP s) {
    if (Rf_isNull(s)) return s;
    SEXP d = Rf_allocVector(s->type, s->len);
    size_t el = s->type == REALSXP ? 8 : (s->type == INTSXP || s->type == LGLSXP) ? 4 : sizeof(void *);
    memcpy(d->data, s->data, el * (size_t)s->len);
    d->dim = s->dim;
    d->names = s->names;
    return d;
}
SEXP Rf_getAttrib(SEXP s, SEXP what) {
    if (what == R_NamesSymbol) return s->names ? s->names : R_NilValue;
    if (what == R_DimSymbol) return s->dim ? s->dim : R_NilValue;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP what, SEXP v) {
    if (what == R_NamesSymbol) s->names = v;
    if (what == R_DimSymbol) s->dim = v;
    return v;
}
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int LENGTH(SEXP s) { return (int)s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
void Rprintf(const char *fmt, ...) {
    if (g_quiet) return;
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
void REprintf(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
void Rf_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    throw shim_r_error(buf);
}
void Rf_warning(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
int R_finite(double x) { return std::isfinite(x); }
Rboolean R_ToplevelExec(void (*fun)(void *), void *data) { fun(data); return TRUE; }
void R_CheckUserInterrupt(void) {}
DL_FUNC R_GetCCallable(const char *, const char *) { return nullptr; }

// ------------------------------------------------------------------ GDS model
struct MemGeno {
    const uint8_t *data;   // [nsnp][nsamp], SNP-major, sample fastest ("sample.order")
    int nsnp, nsamp;
};
static std::string g_err;

int GDS_Array_DimCnt(PdAbstractArray) { return 2; }
void GDS_Array_GetDim(PdAbstractArray obj, C_Int32 *out, int n) {
    MemGeno *g = (MemGeno *)obj;
    if (n >= 2) { out[0] = g->nsnp; out[1] = g->nsamp; }
}
void GDS_Array_ReadDataEx(PdAbstractArray obj, const C_Int32 *st, const C_Int32 *len,
                          const C_BOOL *const sel[], void *out, enum C_SVType sv) {
    MemGeno *g = (MemGeno *)obj;
    if (sv != svUInt8) throw ErrCoreArray("shim: only svUInt8 reads are modelled");
    uint8_t *o = (uint8_t *)out;
    for (int r = 0; r < len[0]; r++) {
        if (sel && sel[0] && !sel[0][r]) continue;
        const uint8_t *row = g->data + (size_t)(st[0] + r) * g->nsamp + st[1];
        if (sel && sel[1]) {
            for (int c = 0; c < len[1]; c++)
                if (sel[1][c]) *o++ = row[c];
        } else {
            memcpy(o, row, (size_t)len[1]);
            o += len[1];
        }
    }
}
void GDS_Array_ReadData(PdAbstractArray obj, const C_Int32 *st, const C_Int32 *len, void *out,
                        enum C_SVType sv) {
    GDS_Array_ReadDataEx(obj, st, len, nullptr, out, sv);
}
// GDS output model: ONE extendable float64 node ("grm" of a SNPRELATE_OUTPUT file, R/IBD.R:589-590);
// appended values are kept in memory in arrival order so that a test can read the stream back
static std::vector<double> g_gds_out;
static int g_gds_out_node = 0;            // its address stands for the node
void GDS_Array_AppendData(PdAbstractArray obj, ssize_t cnt, const void *in, enum C_SVType sv) {
    if (obj != (PdAbstractArray)&g_gds_out_node) throw ErrCoreArray("shim: append to an unknown GDS node");
    if (sv != svFloat64) throw ErrCoreArray("shim: only svFloat64 appends are modelled");
    const double *p = (const double *)in;
    g_gds_out.insert(g_gds_out.end(), p, p + cnt);
}
int GDS_Attr_Name2Index(PdGDSObj, const char *name) { return strcmp(name, "sample.order") == 0 ? 0 : -1; }
static int g_gds_out_root = 0;
PdGDSObj GDS_Node_Path(PdGDSFolder root, const char *path, C_BOOL) {
    if (root == (PdGDSFolder)&g_gds_out_root && strcmp(path, "grm") == 0) return (PdGDSObj)&g_gds_out_node;
    throw ErrCoreArray("shim: no GDS file");
}
PdGDSObj GDS_R_SEXP2Obj(SEXP, C_BOOL) { throw ErrCoreArray("shim: no GDS file"); }
static shim_sexp g_gds_out_file{VECSXP, 0, nullptr, nullptr, nullptr};     // stands for the R object of the output file
PdGDSFolder GDS_R_SEXP2FileRoot(SEXP s) {
    if (s == &g_gds_out_file) return (PdGDSFolder)&g_gds_out_root;
    throw ErrCoreArray("shim: no GDS file");
}
void GDS_Iter_GetStart(PdAbstractArray, PdIterator) { throw ErrCoreArray("shim: iterators are not modelled"); }
C_Float64 GDS_Iter_GetFloat(PdIterator) { throw ErrCoreArray("shim: iterators are not modelled"); }
C_UInt64 GDS_Mach_GetCPULevelCache(int level) {
    long v = 0;
#ifdef _SC_LEVEL1_DCACHE_SIZE
    if (level == 1) v = sysconf(_SC_LEVEL1_DCACHE_SIZE);
    if (level == 2) v = sysconf(_SC_LEVEL2_CACHE_SIZE);
    if (level == 3) v = sysconf(_SC_LEVEL3_CACHE_SIZE);
#endif
    if (v <= 0) v = level == 1 ? 32 * 1024 : level == 2 ? 1024 * 1024 : 0;
    return (C_UInt64)v;
}
PdThreadMutex GDS_Parallel_InitMutex(void) {
    pthread_mutex_t *m = new pthread_mutex_t;
    pthread_mutex_init(m, nullptr);
    return m;
}
void GDS_Parallel_DoneMutex(PdThreadMutex m) {
    if (m) { pthread_mutex_destroy((pthread_mutex_t *)m); delete (pthread_mutex_t *)m; }
}
void GDS_Parallel_LockMutex(PdThreadMutex m) { if (m) pthread_mutex_lock((pthread_mutex_t *)m); }
void GDS_Parallel_UnlockMutex(PdThreadMutex m) { if (m) pthread_mutex_unlock((pthread_mutex_t *)m); }
struct Suspend { pthread_mutex_t m; pthread_cond_t c; };
PdThreadsSuspending GDS_Parallel_InitSuspend(void) {
    Suspend *s = new Suspend;
    pthread_mutex_init(&s->m, nullptr);
    pthread_cond_init(&s->c, nullptr);
    return s;
}
void GDS_Parallel_DoneSuspend(PdThreadsSuspending p) {
    Suspend *s = (Suspend *)p;
    if (s) { pthread_cond_destroy(&s->c); pthread_mutex_destroy(&s->m); delete s; }
}
void GDS_Parallel_Suspend(PdThreadsSuspending p) {
    Suspend *s = (Suspend *)p;
    pthread_mutex_lock(&s->m);
    pthread_cond_wait(&s->c, &s->m);
    pthread_mutex_unlock(&s->m);
}
void GDS_Parallel_WakeUp(PdThreadsSuspending p) {
    Suspend *s = (Suspend *)p;
    pthread_mutex_lock(&s->m);
    pthread_cond_broadcast(&s->c);
    pthread_mutex_unlock(&s->m);
}
struct RunArg { void (*proc)(PdThread, int, void *); void *param; int idx; };
static void *run_thread(void *p) {
    RunArg *a = (RunArg *)p;
    a->proc(nullptr, a->idx, a->param);
    return nullptr;
}
void GDS_Parallel_RunThreads(void (*proc)(PdThread, int, void *), void *param, int nthread) {
    if (nthread <= 1) { proc(nullptr, 0, param); return; }
    std::vector<pthread_t> th(nthread - 1);
    std::vector<RunArg> args(nthread);
    for (int i = 0; i < nthread; i++) args[i] = RunArg{proc, param, i};
    for (int i = 1; i < nthread; i++) pthread_create(&th[i - 1], nullptr, run_thread, &args[i]);
    proc(nullptr, 0, param);
    for (int i = 1; i < nthread; i++) pthread_join(th[i - 1], nullptr);
}
void GDS_SetError(const char *msg) { g_err = msg ? msg : ""; }
const char *GDS_GetError(void) { return g_err.c_str(); }
}  // extern "C"

// ------------------------------------------------------------------ driver
extern "C" {
SEXP gnrGRM(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrPCA(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrEigMix(SEXP, SEXP, SEXP, SEXP);
SEXP gnrIBSAve(SEXP, SEXP, SEXP);
SEXP gnrIBSNum(SEXP, SEXP);
SEXP gnrIBD_KING_Robust(SEXP, SEXP, SEXP, SEXP);
SEXP gnrIBD_KING_Homo(SEXP, SEXP, SEXP);
SEXP gnrIBD_PLINK(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrIBD_Beta(SEXP, SEXP, SEXP, SEXP);
SEXP gnrPCACorr(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrPCASNPLoading(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrPCASampLoading(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrEigMixSNPLoading(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP gnrEigMixSampLoading(SEXP, SEXP, SEXP, SEXP);
}

static MemGeno g_geno;
static std::vector<uint8_t> g_store;

#define REF_TRY try {
#define REF_CATCH                                                      \
    return 0;                                                          \
    }                                                                  \
    catch (std::exception & e) { g_err = e.what(); return 1; }         \
    catch (const char *e) { g_err = e; return 1; }                     \
    catch (...) { g_err = "unknown error"; return 1; }

static SEXP named_list(std::initializer_list<std::pair<const char *, SEXP>> items) {
    SEXP l = Rf_allocVector(VECSXP, (R_xlen_t)items.size());
    SEXP nm = Rf_allocVector(STRSXP, (R_xlen_t)items.size());
    R_xlen_t i = 0;
    for (auto &kv : items) {
        SET_VECTOR_ELT(l, i, kv.second);
        ((SEXP *)nm->data)[i] = Rf_mkChar(kv.first);
        i++;
    }
    l->names = nm;
    return l;
}

extern "C" {
const char *ref_error(void) { return g_err.c_str(); }
void ref_verbose(int v) { g_quiet = !v; }

// gnrSetGenoSpace (src/SNPRelate.cpp:76-114) on an in-memory SNP-major matrix (copied)
int ref_set_geno(const uint8_t *geno, int nsnp, int nsamp) {
    REF_TRY
    g_store.assign(geno, geno + (size_t)nsnp * nsamp);
    g_geno = MemGeno{g_store.data(), nsnp, nsamp};
    GWAS::MCWorkingGeno.InitSNPGDSFile(&g_geno, false);
    GWAS::MCWorkingGeno.Space().InitSelection();
    if (GWAS::MCWorkingGeno.Space().SNPNum() <= 0) throw ErrCoreArray("There is no SNP!");
    if (GWAS::MCWorkingGeno.Space().SampleNum() <= 0) throw ErrCoreArray("There is no sample!");
    REF_CATCH
}
// gnrSelSNP_Base (src/SNPRelate.cpp:184-210)
int ref_select_snp_base(int remove_mono, double maf, double missrate, uint8_t *flags, int *n_removed) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SNPNum();
    std::vector<C_BOOL> sel(n);
    int out = GWAS::MCWorkingGeno.Space().Select_SNP_Base(remove_mono != 0, maf, missrate, &sel[0]);
    if (flags) for (int i = 0; i < n; i++) flags[i] = sel[i] ? 1 : 0;
    if (n_removed) *n_removed = out;
    REF_CATCH
}
int ref_dims(int *nsnp, int *nsamp) {
    REF_TRY
    *nsnp = GWAS::MCWorkingGeno.Space().SNPNum();
    *nsamp = GWAS::MCWorkingGeno.Space().SampleNum();
    REF_CATCH
}
static void copy_real(SEXP s, double *out) { memcpy(out, REAL(s), sizeof(double) * (size_t)s->len); }

int ref_grm(const char *method, int nthread, double *out) {
    REF_TRY
    SEXP r = gnrGRM(Rf_ScalarInteger(nthread), Rf_mkString(method), R_NilValue, Rf_ScalarLogical(0),
                    Rf_ScalarLogical(0));
    copy_real(r, out);
    REF_CATCH
}
// gnrGRM with a GDS output file (R/IBD.R:570-594): returns the stream appended to the "grm" node
int ref_grm_gds(const char *method, int nthread, double *out, long long *count) {
    REF_TRY
    g_gds_out.clear();
    gnrGRM(Rf_ScalarInteger(nthread), Rf_mkString(method), &g_gds_out_file, Rf_ScalarLogical(1), Rf_ScalarLogical(0));
    if (count) *count = (long long)g_gds_out.size();
    if (out) memcpy(out, g_gds_out.data(), g_gds_out.size() * sizeof(double));
    REF_CATCH
}
int ref_pca(int nthread, int bayesian, int eigen_cnt, double *genmat, double *trace_xtx, double *eigval,
            double *eigvec) {
    REF_TRY
    SEXP param = named_list({{"bayesian", Rf_ScalarLogical(bayesian)},
                             {"need.genmat", Rf_ScalarLogical(genmat != nullptr)},
                             {"genmat.only", Rf_ScalarLogical(eigen_cnt <= 0)},
                             {"eigen.method", Rf_mkString("DSPEVX")}});
    SEXP r = gnrPCA(Rf_ScalarInteger(eigen_cnt > 0 ? eigen_cnt : 1), Rf_mkString("exact"),
                    Rf_ScalarInteger(nthread), param, Rf_ScalarLogical(0));
    if (trace_xtx) *trace_xtx = REAL(VECTOR_ELT(r, 0))[0];
    if (genmat) copy_real(VECTOR_ELT(r, 1), genmat);
    if (eigen_cnt > 0) {
        if (eigval) copy_real(VECTOR_ELT(r, 2), eigval);
        if (eigvec) copy_real(VECTOR_ELT(r, 3), eigvec);
    }
    REF_CATCH
}
// gnrPCA "randomized" (src/genPCA.cpp:1436-1442): aux.mat [aux.dim * n.samp] is overwritten
int ref_pca_randomized(int nthread, double *aux_mat, int aux_dim, int iter_num, double *sigma, double *vt,
                       double *trace2) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SampleNum();
    SEXP aux = Rf_allocVector(REALSXP, (R_xlen_t)aux_dim * n);
    memcpy(REAL(aux), aux_mat, sizeof(double) * (size_t)aux_dim * n);
    SEXP param = named_list({{"aux.mat", aux}, {"aux.dim", Rf_ScalarInteger(aux_dim)},
                             {"iter.num", Rf_ScalarInteger(iter_num)}});
    SEXP r = gnrPCA(Rf_ScalarInteger(1), Rf_mkString("randomized"), Rf_ScalarInteger(nthread), param,
                    Rf_ScalarLogical(0));
    if (sigma) copy_real(VECTOR_ELT(r, 0), sigma);
    if (vt) copy_real(VECTOR_ELT(r, 1), vt);
    if (trace2) *trace2 = REAL(VECTOR_ELT(r, 2))[0];
    REF_CATCH
}
int ref_eigmix(int nthread, int diagadj, double *ibd, double *afreq) {
    REF_TRY
    SEXP param = named_list({{"diagadj", Rf_ScalarLogical(diagadj)}, {"ibdmat", Rf_ScalarLogical(1)}});
    SEXP r = gnrEigMix(Rf_ScalarInteger(0), Rf_ScalarInteger(nthread), param, Rf_ScalarLogical(0));
    if (afreq) copy_real(VECTOR_ELT(r, 2), afreq);
    if (ibd) copy_real(VECTOR_ELT(r, 3), ibd);
    REF_CATCH
}
int ref_ibs_num(int nthread, int *i0, int *i1, int *i2) {
    REF_TRY
    SEXP r = gnrIBSNum(Rf_ScalarInteger(nthread), Rf_ScalarLogical(0));
    int *o[3] = {i0, i1, i2};
    for (int k = 0; k < 3; k++) {
        SEXP m = VECTOR_ELT(r, k);
        memcpy(o[k], INTEGER(m), sizeof(int) * (size_t)m->len);
    }
    REF_CATCH
}
int ref_ibs_ave(int nthread, double *out) {
    REF_TRY
    copy_real(gnrIBSAve(Rf_ScalarInteger(nthread), Rf_ScalarLogical(0), Rf_ScalarLogical(0)), out);
    REF_CATCH
}
int ref_king_robust(int nthread, const int *family, double *ibs0, double *kinship) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SampleNum();
    SEXP fam = Rf_allocVector(INTSXP, n);
    for (int i = 0; i < n; i++) INTEGER(fam)[i] = family ? family[i] : R_NaInt;
    SEXP r = gnrIBD_KING_Robust(fam, Rf_ScalarInteger(nthread), Rf_ScalarLogical(0), Rf_ScalarLogical(0));
    copy_real(VECTOR_ELT(r, 0), ibs0);
    copy_real(VECTOR_ELT(r, 1), kinship);
    REF_CATCH
}
int ref_king_homo(int nthread, double *k0, double *k1) {
    REF_TRY
    SEXP r = gnrIBD_KING_Homo(Rf_ScalarInteger(nthread), Rf_ScalarLogical(0), Rf_ScalarLogical(0));
    copy_real(VECTOR_ELT(r, 0), k0);
    copy_real(VECTOR_ELT(r, 1), k1);
    REF_CATCH
}
int ref_ibd_mom(int nthread, const double *afreq_in, int constraint, double *k0, double *k1, double *afreq) {
    REF_TRY
    const int m = GWAS::MCWorkingGeno.Space().SNPNum();
    SEXP af = Rf_allocVector(REALSXP, afreq_in ? m : 0);
    if (afreq_in) memcpy(REAL(af), afreq_in, sizeof(double) * (size_t)m);
    SEXP r = gnrIBD_PLINK(Rf_ScalarInteger(nthread), af, Rf_ScalarLogical(afreq_in != nullptr),
                          Rf_ScalarLogical(constraint), Rf_ScalarLogical(0), Rf_ScalarLogical(0));
    copy_real(VECTOR_ELT(r, 0), k0);
    copy_real(VECTOR_ELT(r, 1), k1);
    copy_real(VECTOR_ELT(r, 2), afreq);
    REF_CATCH
}
static SEXP real_matrix(const double *src, int nr, int nc) {
    SEXP m = Rf_allocMatrix(REALSXP, nr, nc);
    memcpy(REAL(m), src, sizeof(double) * (size_t)nr * nc);
    return m;
}
static SEXP real_vector(const double *src, int n) {
    SEXP v = Rf_allocVector(REALSXP, n);
    memcpy(REAL(v), src, sizeof(double) * (size_t)n);
    return v;
}
// gnrPCACorr (src/genPCA.cpp:1456-1485): eigvect n x k column-major -> corr k x nsnp
int ref_pca_corr(int nthread, int k, const double *eigvect, double *out) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SampleNum();
    copy_real(gnrPCACorr(Rf_ScalarInteger(k), real_matrix(eigvect, n, k), Rf_ScalarInteger(nthread), R_NilValue,
                         Rf_ScalarLogical(0)), out);
    REF_CATCH
}
// gnrPCASNPLoading (src/genPCA.cpp:1489-1540) -> loading k x nsnp, avgfreq, scale
int ref_pca_snp_loading(int nthread, int k, const double *eigval, const double *eigvect, double trace_xtx,
                        int bayesian, double *loading, double *avgfreq, double *scale) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SampleNum();
    SEXP r = gnrPCASNPLoading(real_vector(eigval, k), real_matrix(eigvect, n, k), Rf_ScalarReal(trace_xtx),
                              Rf_ScalarInteger(nthread), Rf_ScalarLogical(bayesian), Rf_ScalarLogical(0));
    copy_real(VECTOR_ELT(r, 0), loading);
    copy_real(VECTOR_ELT(r, 1), avgfreq);
    copy_real(VECTOR_ELT(r, 2), scale);
    REF_CATCH
}
// gnrPCASampLoading (src/genPCA.cpp:1542-1563): loadings k x nsnp -> n x k
int ref_pca_samp_loading(int nthread, int k, const double *loadings, const double *avgfreq, const double *scale,
                         double *out) {
    REF_TRY
    const int m = GWAS::MCWorkingGeno.Space().SNPNum();
    copy_real(gnrPCASampLoading(Rf_ScalarInteger(k), real_matrix(loadings, k, m), real_vector(avgfreq, m),
                                real_vector(scale, m), Rf_ScalarInteger(nthread), Rf_ScalarLogical(0)), out);
    REF_CATCH
}
// gnrEigMixSNPLoading (src/genEIGMIX.cpp:739-775)
int ref_eigmix_snp_loading(int nthread, int k, const double *eigval, const double *eigvect, const double *afreq,
                           double *loading) {
    REF_TRY
    const int n = GWAS::MCWorkingGeno.Space().SampleNum(), m = GWAS::MCWorkingGeno.Space().SNPNum();
    copy_real(gnrEigMixSNPLoading(real_vector(eigval, k), real_matrix(eigvect, n, k), real_vector(afreq, m),
                                  Rf_ScalarInteger(nthread), Rf_ScalarLogical(0)), loading);
    REF_CATCH
}
// gnrEigMixSampLoading (src/genEIGMIX.cpp:777-803)
int ref_eigmix_samp_loading(int nthread, int k, const double *loadings, const double *afreq, double *out) {
    REF_TRY
    const int m = GWAS::MCWorkingGeno.Space().SNPNum();
    copy_real(gnrEigMixSampLoading(real_matrix(loadings, k, m), real_vector(afreq, m), Rf_ScalarInteger(nthread),
                                   Rf_ScalarLogical(0)), out);
    REF_CATCH
}
int ref_indiv_beta(int nthread, int inbreeding, double *out) {
    REF_TRY
    copy_real(gnrIBD_Beta(Rf_ScalarLogical(inbreeding), Rf_ScalarInteger(nthread), Rf_ScalarLogical(0),
                          Rf_ScalarLogical(0)), out);
    REF_CATCH
}
}  // extern "C"

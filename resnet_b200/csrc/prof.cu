// prof.cu -- see prof.h
#include "prof.h"
#include <vector>

namespace rb {
struct Rec { cudaEvent_t a, b; int fam; double work, work2; };
static bool g_on = false;
static std::vector<Rec> g_recs;
static size_t g_used = 0;

void prof_enable(bool on) { g_on = on; }
bool prof_enabled() { return g_on; }
void prof_begin(cudaStream_t st, int family, double work, double work2) {
	if (g_used == g_recs.size()) {
		Rec r;
		RB_CUDA(cudaEventCreate(&r.a));
		RB_CUDA(cudaEventCreate(&r.b));
		g_recs.push_back(r);
	}
	Rec &r = g_recs[g_used];
	r.fam = family;
	r.work = work;
	r.work2 = work2;
	RB_CUDA(cudaEventRecord(r.a, st));
}
void prof_end(cudaStream_t st) {
	RB_CUDA(cudaEventRecord(g_recs[g_used].b, st));
	g_used++;
}
int prof_read(int family, double *ms, long long *launches, double *work, double *work2) {
	double t = 0, w = 0, w2 = 0;
	long long n = 0;
	for (size_t i = 0; i < g_used; i++) {
		if (g_recs[i].fam != family) continue;
		RB_CUDA(cudaEventSynchronize(g_recs[i].b));
		float e = 0;
		RB_CUDA(cudaEventElapsedTime(&e, g_recs[i].a, g_recs[i].b));
		t += e;
		w += g_recs[i].work;
		w2 += g_recs[i].work2;
		n++;
	}
	*ms = t; *launches = n; *work = w;
	if (work2) *work2 = w2;
	return has_error() ? 1 : 0;
}
void prof_reset() { g_used = 0; }
}  // namespace rb

// Symbolic analysis of one QP sparsity pattern (host, integers only, once per pattern).
// See plan.h for the list of reference functions this replaces.

#include "plan.h"

#include <cstring>

namespace qpb {

uint64_t fnv1a_words(const long *a, long count, uint64_t h)
{
	for (long i = 0; i < count; ++i) {
		h ^= (uint64_t)a[i];
		h *= 1099511628211ULL;
	}
	return h;
}

long Plan::factor_flops() const
{
	long f = 0;
	for (long k = 0; k < N; ++k) f += Lnz[k] * (Lnz[k] + 2);
	return f;
}

namespace {

// Row-wise view of a CCS pattern: for every row the (column, stored position) pairs in
// ascending column order.  This is the order in which the reference's column-scatter
// products (SparseMatrixMultiply, Auxilary.c:853-859) touch an output row, and also the
// order of the entries inside a column of the transposes At/Gt (Auxilary.c:937-947).
struct RowView
{
	std::vector<int32_t> ptr, col, pos;
};

RowView rows_of(long nrows, long ncols, const long *jc, const long *ir)
{
	RowView v;
	v.ptr.assign(nrows + 1, 0);
	const long nz = jc ? jc[ncols] : 0;
	v.col.assign(nz, 0);
	v.pos.assign(nz, 0);
	for (long q = 0; q < nz; ++q) ++v.ptr[ir[q] + 1];
	for (long r = 0; r < nrows; ++r) v.ptr[r + 1] += v.ptr[r];
	std::vector<int32_t> fill(v.ptr.begin(), v.ptr.end() - 1);
	for (long c = 0; c < ncols; ++c)
		for (long q = jc[c]; q < jc[c + 1]; ++q) {
			const int32_t at = fill[ir[q]]++;
			v.col[at] = (int32_t)c;
			v.pos[at] = (int32_t)q;
		}
	return v;
}

template <class T>
int32_t append(std::vector<int32_t> &blob, const std::vector<T> &section)
{
	const int32_t at = (int32_t)blob.size();
	for (size_t i = 0; i < section.size(); ++i) blob.push_back((int32_t)section[i]);
	return at;
}

} // namespace

bool build_plan(Plan &plan, long n, long m, long p, const long *Pjc, const long *Pir, const long *Ajc, const long *Air, const long *Gjc, const long *Gir, const long *user_perm)
{
	if (!Ajc || !Air) p = 0;
	plan.n = n;
	plan.m = m;
	plan.p = p;
	const long N = n + p + m;
	plan.N = N;
	plan.nnzP = Pjc[n];
	plan.nnzA = p ? Ajc[n] : 0;
	plan.nnzG = Gjc[n];
	const long nV = plan.nnzP + plan.nnzA + plan.nnzG;
	const long offA = plan.nnzP, offG = plan.nnzP + plan.nnzA;

	const RowView prow = rows_of(n, n, Pjc, Pir);
	const RowView arow = p ? rows_of(p, n, Ajc, Air) : RowView{std::vector<int32_t>(1, 0), {}, {}};
	const RowView grow = rows_of(m, n, Gjc, Gir);

	// ---- KKT pattern, column by column, in the reference's storage order -----------------
	// (Auxilary.c:82-134 with equalities, 145-179 without).  The (z,z) diagonal is emitted right
	// after the LAST entry of the matching G row, so an empty G row would silently lose its
	// diagonal in the reference; we refuse such patterns instead of corrupting memory.
	plan.valid = true;
	for (long r = 0; r < m; ++r)
		if (grow.ptr[r + 1] == grow.ptr[r]) plan.valid = false;
	plan.Kjc.assign(N + 1, 0);
	plan.Kir.clear();
	plan.Ksrc.clear();
	for (long c = 0; c < n; ++c) {
		for (long q = Pjc[c]; q < Pjc[c + 1]; ++q) { plan.Kir.push_back(Pir[q]); plan.Ksrc.push_back((int32_t)q); }
		if (p)
			for (long q = Ajc[c]; q < Ajc[c + 1]; ++q) { plan.Kir.push_back(n + Air[q]); plan.Ksrc.push_back((int32_t)(offA + q)); }
		for (long q = Gjc[c]; q < Gjc[c + 1]; ++q) { plan.Kir.push_back(n + p + Gir[q]); plan.Ksrc.push_back((int32_t)(offG + q)); }
		plan.Kjc[c + 1] = (long)plan.Kir.size();
	}
	for (long r = 0; r < p; ++r) {
		for (int32_t t = arow.ptr[r]; t < arow.ptr[r + 1]; ++t) { plan.Kir.push_back(arow.col[t]); plan.Ksrc.push_back((int32_t)(offA + arow.pos[t])); }
		plan.Kjc[n + r + 1] = (long)plan.Kir.size();
	}
	for (long r = 0; r < m; ++r) {
		for (int32_t t = grow.ptr[r]; t < grow.ptr[r + 1]; ++t) { plan.Kir.push_back(grow.col[t]); plan.Ksrc.push_back((int32_t)(offG + grow.pos[t])); }
		if (grow.ptr[r + 1] > grow.ptr[r]) { plan.Kir.push_back(n + p + r); plan.Ksrc.push_back((int32_t)(nV + r)); }
		plan.Kjc[n + p + r + 1] = (long)plan.Kir.size();
	}

	// ---- ordering (qpSWIFT.c:194-219) ----------------------------------------------------------
	plan.P.assign(N, 0);
	if (user_perm) {
		for (long k = 0; k < N; ++k) plan.P[k] = user_perm[k];
		plan.amd_status = -3;
	} else {
		plan.amd_status = amd_order(N, plan.Kjc.data(), plan.Kir.data(), plan.P.data());
		if (plan.amd_status < 0)
			for (long k = 0; k < N; ++k) plan.P[k] = k;
	}

	// ---- elimination tree and column counts (ldl.c:187-240) ------------------------------------
	plan.Pinv.assign(N, 0);
	plan.Parent.assign(N, -1);
	plan.Lnz.assign(N, 0);
	plan.Lp.assign(N + 1, 0);
	for (long k = 0; k < N; ++k) plan.Pinv[plan.P[k]] = k;
	std::vector<long> visited(N, -1);
	for (long k = 0; k < N; ++k) {
		visited[k] = k;
		const long col = plan.P[k];
		for (long q = plan.Kjc[col]; q < plan.Kjc[col + 1]; ++q) {
			long i = plan.Pinv[plan.Kir[q]];
			if (i >= k) continue;
			while (visited[i] != k) {
				if (plan.Parent[i] == -1) plan.Parent[i] = k;
				++plan.Lnz[i];
				visited[i] = k;
				i = plan.Parent[i];
			}
		}
	}
	for (long k = 0; k < N; ++k) plan.Lp[k + 1] = plan.Lp[k] + plan.Lnz[k];
	const long lnz = plan.Lp[N];

	// ---- integer half of LDL_numeric (ldl.c:276-298, 315): row patterns, Li, store slots ---------
	plan.Li.assign(lnz, 0);
	plan.pat_ptr.assign(N + 1, 0);
	plan.pat_col.clear();
	plan.pat_slot.clear();
	plan.scat_ptr.assign(N + 1, 0);
	plan.scat_dst.clear();
	plan.scat_src.clear();
	std::vector<long> filled(N, 0), stack(N, 0), path(N, 0);
	std::fill(visited.begin(), visited.end(), -1);
	for (long k = 0; k < N; ++k) {
		long top = N;
		visited[k] = k;
		const long col = plan.P[k];
		for (long q = plan.Kjc[col]; q < plan.Kjc[col + 1]; ++q) {
			long i = plan.Pinv[plan.Kir[q]];
			if (i > k) continue;
			plan.scat_dst.push_back((int32_t)i);
			plan.scat_src.push_back(plan.Ksrc[q]);
			long len = 0;
			for (; visited[i] != k; i = plan.Parent[i]) { path[len++] = i; visited[i] = k; }
			while (len > 0) stack[--top] = path[--len];
		}
		plan.scat_ptr[k + 1] = (int32_t)plan.scat_dst.size();
		for (; top < N; ++top) {
			const long i = stack[top];
			const long slot = plan.Lp[i] + filled[i];
			plan.Li[slot] = k;
			++filled[i];
			plan.pat_col.push_back((int32_t)i);
			plan.pat_slot.push_back((int32_t)slot);
		}
		plan.pat_ptr[k + 1] = (int32_t)plan.pat_col.size();
	}

	// ---- Transpose_Row_Count (Auxilary.c:683-732): kept for struct compatibility ------------------
	plan.Ltp.assign(N + 1, 0);
	plan.Lti.assign(lnz > 0 ? lnz : 1, 0);
	{
		std::vector<long> cnt(N, 0);
		for (long q = 0; q < lnz; ++q) ++cnt[plan.Li[q]];
		for (long r = 0; r < N; ++r) plan.Ltp[r + 1] = plan.Ltp[r] + cnt[r];
		std::fill(cnt.begin(), cnt.end(), 0);
		for (long c = 0; c < N; ++c)
			for (long q = plan.Lp[c]; q < plan.Lp[c + 1]; ++q) plan.Lti[plan.Ltp[plan.Li[q]] + cnt[plan.Li[q]]++] = c;
	}

	// ---- pattern identity ---------------------------------------------------------------------
	{
		long dims[4] = {n, m, p, user_perm ? 1 : 0};
		uint64_t h = fnv1a_words(dims, 4);
		h = fnv1a_words(Pjc, n + 1, h);
		h = fnv1a_words(Pir, plan.nnzP, h);
		if (p) { h = fnv1a_words(Ajc, n + 1, h); h = fnv1a_words(Air, plan.nnzA, h); }
		h = fnv1a_words(Gjc, n + 1, h);
		h = fnv1a_words(Gir, plan.nnzG, h);
		h = fnv1a_words(plan.P.data(), N, h);
		plan.pattern_hash = h;
	}

	// ---- device image -------------------------------------------------------------------------
	DevicePlanHeader hd;
	std::memset(&hd, 0, sizeof hd);
	hd.n = (int32_t)n; hd.m = (int32_t)m; hd.p = (int32_t)p; hd.N = (int32_t)N;
	hd.nnzP = (int32_t)plan.nnzP; hd.nnzA = (int32_t)plan.nnzA; hd.nnzG = (int32_t)plan.nnzG; hd.nV = (int32_t)nV;
	hd.lnz = (int32_t)lnz;
	hd.nscat = (int32_t)plan.scat_dst.size();
	std::vector<int32_t> &blob = plan.blob;
	blob.assign(sizeof(DevicePlanHeader) / sizeof(int32_t), 0);
	hd.perm = append(blob, plan.P);
	hd.lp = append(blob, plan.Lp);
	hd.li = append(blob, plan.Li);
	hd.scat_ptr = append(blob, plan.scat_ptr);
	hd.scat_dst = append(blob, plan.scat_dst);
	hd.scat_src = append(blob, plan.scat_src);
	hd.pat_ptr = append(blob, plan.pat_ptr);
	hd.pat_col = append(blob, plan.pat_col);
	hd.pat_slot = append(blob, plan.pat_slot);
	hd.prow_ptr = append(blob, prow.ptr);
	hd.prow_col = append(blob, prow.col);
	hd.prow_val = append(blob, prow.pos);
	{
		std::vector<int32_t> v(arow.pos);
		for (size_t i = 0; i < v.size(); ++i) v[i] += (int32_t)offA;
		hd.arow_ptr = append(blob, arow.ptr);
		hd.arow_col = append(blob, arow.col);
		hd.arow_val = append(blob, v);
	}
	{
		std::vector<int32_t> v(grow.pos);
		for (size_t i = 0; i < v.size(); ++i) v[i] += (int32_t)offG;
		hd.grow_ptr = append(blob, grow.ptr);
		hd.grow_col = append(blob, grow.col);
		hd.grow_val = append(blob, v);
	}
	{
		std::vector<long> aptr(n + 1, 0), arows;
		if (p) { aptr.assign(Ajc, Ajc + n + 1); arows.assign(Air, Air + plan.nnzA); }
		hd.acol_ptr = append(blob, aptr);
		hd.acol_row = append(blob, arows);
		std::vector<long> gptr(Gjc, Gjc + n + 1), grows(Gir, Gir + plan.nnzG);
		hd.gcol_ptr = append(blob, gptr);
		hd.gcol_row = append(blob, grows);
	}
	hd.words = (int32_t)blob.size();
	std::memcpy(blob.data(), &hd, sizeof hd);
	return plan.valid;
}

} // namespace qpb

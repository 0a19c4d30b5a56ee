// Approximate-minimum-degree ordering, written from scratch for the B200 qpSWIFT plan builder.
//
// The symbolic plan (permutation -> elimination tree -> factor pattern -> static kernel
// schedule) must be BIT-EXACT with what the reference's embedded SuiteSparse AMD 2.3.1
// produces, because the permutation fixes the floating-point summation order of everything
// that runs on the GPU afterwards.  "A good ordering" is therefore not enough: this file
// reproduces every tie-break of the reference pipeline
//
//   amd_l_order  (src/qpSWIFT/amd_order.c:21-199)    driver: validate, [transpose], count, order
//   amd_l_valid  (src/qpSWIFT/amd_valid.c:38-92)     sorted/duplicate-free check
//   amd_l_preprocess (src/qpSWIFT/amd_preprocess.c:29-118)  R = pattern(A') without duplicates
//   amd_l_aat    (src/qpSWIFT/amd_aat.c:20-184)      degrees of A+A'
//   amd_l1       (src/qpSWIFT/amd_1.c:29-180)        adjacency lists of A+A'
//   amd_l2       (src/qpSWIFT/amd_2.c:42-1842)       quotient-graph elimination
//   amd_l_postorder / amd_l_post_tree (src/qpSWIFT/amd_postorder.c:15-206, amd_post_tree.c:15-120)
//
// with the reference's defaults (dense = 10, aggressive absorption on, amd_defaults.c:21-37).
// The code is organised differently (one quotient-graph object, one method per phase, 0-based
// std::vector storage) but the list insertion orders, the unsigned hash, the LIFO degree lists,
// the ">=" tie-break of the post-order and the workspace size that decides when the adjacency
// store is compacted are the reference's; see SURVEY.md section 8a "A2 tie-breaks".
//
// Host-side integer work, run once per sparsity pattern.  No CUDA here.

#include "plan.h"

#include <cmath>
#include <cstddef>
#include <vector>

namespace qpb {

namespace {

typedef long idx;
const idx NONE = -1;
// Involution used to mark "this slot stores a parent / a hash head / a flipped count".
inline idx mirror(idx v) { return -v - 2; }

// ---------------------------------------------------------------------------------------
// Structure checks and the optional transpose used for unsorted input.
// ---------------------------------------------------------------------------------------

// 0 = sorted and duplicate free, 1 = usable but unsorted or with duplicates, -2 = invalid.
int classify_columns(idx n, const idx *colptr, const idx *rowidx)
{
	if (n < 0 || !colptr || !rowidx) return AMD_STATUS_INVALID;
	if (colptr[0] != 0 || colptr[n] < 0) return AMD_STATUS_INVALID;
	int verdict = AMD_STATUS_OK;
	for (idx j = 0; j < n; ++j) {
		if (colptr[j] > colptr[j + 1]) return AMD_STATUS_INVALID;
		idx previous = NONE;
		for (idx q = colptr[j]; q < colptr[j + 1]; ++q) {
			const idx r = rowidx[q];
			if (r < 0 || r >= n) return AMD_STATUS_INVALID;
			if (r <= previous) verdict = AMD_STATUS_JUMBLED;
			previous = r;
		}
	}
	return verdict;
}

// Pattern of the transpose with duplicates dropped; columns of the result come out sorted.
void transpose_pattern(idx n, const idx *colptr, const idx *rowidx, std::vector<idx> &tp, std::vector<idx> &ti)
{
	std::vector<idx> fill(n, 0), seen(n, NONE);
	for (idx j = 0; j < n; ++j)
		for (idx q = colptr[j]; q < colptr[j + 1]; ++q) {
			const idx r = rowidx[q];
			if (seen[r] != j) { ++fill[r]; seen[r] = j; }
		}
	tp.assign(n + 1, 0);
	for (idx r = 0; r < n; ++r) tp[r + 1] = tp[r] + fill[r];
	ti.assign(tp[n] > 0 ? tp[n] : 1, 0);
	for (idx r = 0; r < n; ++r) { fill[r] = tp[r]; seen[r] = NONE; }
	for (idx j = 0; j < n; ++j)
		for (idx q = colptr[j]; q < colptr[j + 1]; ++q) {
			const idx r = rowidx[q];
			if (seen[r] != j) { ti[fill[r]++] = j; seen[r] = j; }
		}
}

// ---------------------------------------------------------------------------------------
// One walk over the columns that pairs every strictly-upper entry (j,k) with the not yet
// consumed lower entries of column j.  The same walk is needed twice (first to count the
// degrees of A+A', then to write the adjacency lists), so it is a template over the action
// taken for an undirected edge.  The order in which edges are emitted is what fixes the
// order of every adjacency list, hence it follows amd_aat.c / amd_1.c exactly.
// ---------------------------------------------------------------------------------------
template <class Edge>
void for_each_symmetric_edge(idx n, const idx *colptr, const idx *rowidx, std::vector<idx> &cursor, Edge edge)
{
	cursor.assign(n, 0);
	for (idx k = 0; k < n; ++k) {
		idx q = colptr[k];
		const idx qend = colptr[k + 1];
		while (q < qend) {
			const idx j = rowidx[q];
			if (j > k) break;            // first entry below the diagonal: later
			++q;
			if (j == k) break;           // the diagonal itself is ignored
			edge(j, k);                  // strictly upper entry (j,k)
			// consume the lower part of column j down to row k
			idx t = cursor[j];
			const idx tend = colptr[j + 1];
			while (t < tend) {
				const idx i = rowidx[t];
				if (i > k) break;
				++t;
				if (i == k) break;       // (k,j) mirrors (j,k): already emitted
				edge(i, j);              // entry present only below the diagonal
			}
			cursor[j] = t;
		}
		cursor[k] = q;
	}
	// whatever is left below the diagonals has no mirror image
	for (idx j = 0; j < n; ++j)
		for (idx t = cursor[j]; t < colptr[j + 1]; ++t) edge(rowidx[t], j);
}

// ---------------------------------------------------------------------------------------
// Quotient graph.
// ---------------------------------------------------------------------------------------
struct QuotientGraph
{
	idx n;
	std::vector<idx> store;    // adjacency store: variables' element+variable lists and elements' variable lists
	idx store_cap;             // usable length of `store` (decides when compaction happens)
	idx store_free;            // first free slot
	std::vector<idx> where;    // start of node's list in `store`; mirror(parent) once absorbed; NONE for roots
	std::vector<idx> length;   // length of node's list
	std::vector<idx> weight;   // supervariable size; negated while the variable sits in the current element; 0 = absorbed
	std::vector<idx> nelem;    // #elements at the front of a variable's list; mirror(size) for elements
	std::vector<idx> degree;   // approximate external degree
	std::vector<idx> mark;     // multipurpose stamp array
	std::vector<idx> head;     // degree-list heads, shared with hash-bucket heads (stored mirrored)
	std::vector<idx> fwd;      // next in degree list / hash bucket; final inverse permutation
	std::vector<idx> bwd;      // previous in degree list; hash key during supervariable detection; final permutation

	idx stamp, stamp_limit;
	idx eliminated;
	idx lowest_degree;
	idx widest_element;
	idx dense_cut;
	bool aggressive;

	// current pivot
	idx pivot, pivot_first, pivot_last, pivot_weight, pivot_degree, pivot_old_elems;

	void reset_stamp_if_needed()
	{
		if (stamp < 2 || stamp >= stamp_limit) {
			for (idx v = 0; v < n; ++v)
				if (mark[v] != 0) mark[v] = 1;
			stamp = 2;
		}
	}

	void unlink_from_degree_list(idx v)
	{
		const idx before = bwd[v], after = fwd[v];
		if (after != NONE) bwd[after] = before;
		if (before != NONE) fwd[before] = after;
		else head[degree[v]] = after;
	}

	void push_degree_list(idx v, idx d)
	{
		const idx old_head = head[d];
		if (old_head != NONE) bwd[old_head] = v;
		fwd[v] = old_head;
		head[d] = v;
	}

	void initialise(double dense_ratio, bool aggr)
	{
		aggressive = aggr;
		if (dense_ratio < 0) dense_cut = n - 2;
		else dense_cut = (idx)(dense_ratio * std::sqrt((double)n));
		if (dense_cut < 16) dense_cut = 16;
		if (dense_cut > n) dense_cut = n;

		weight.assign(n, 1);
		nelem.assign(n, 0);
		mark.assign(n, 1);
		head.assign(n, NONE);
		fwd.assign(n, NONE);
		bwd.assign(n, NONE);
		degree = length;
		stamp_limit = __LONG_MAX__ - n;
		stamp = 0;
		reset_stamp_if_needed();
		eliminated = 0;
		lowest_degree = 0;
		widest_element = 0;
		pivot = NONE;

		for (idx v = 0; v < n; ++v) {
			const idx d = degree[v];
			if (d == 0) {                       // isolated: an element right away
				nelem[v] = mirror(1);
				++eliminated;
				where[v] = NONE;
				mark[v] = 0;
			} else if (d > dense_cut) {         // too dense: ordered last, takes no part
				weight[v] = 0;
				nelem[v] = NONE;
				++eliminated;
				where[v] = NONE;
			} else {
				push_degree_list(v, d);
			}
		}
	}

	void choose_pivot()
	{
		idx d = lowest_degree;
		for (; d < n; ++d) {
			pivot = head[d];
			if (pivot != NONE) break;
		}
		lowest_degree = d;
		const idx after = fwd[pivot];
		if (after != NONE) bwd[after] = NONE;
		head[d] = after;
		pivot_old_elems = nelem[pivot];
		pivot_weight = weight[pivot];
		eliminated += pivot_weight;
	}

	// take variable v into the element under construction
	inline void claim(idx v, idx w)
	{
		pivot_degree += w;
		weight[v] = -w;
		unlink_from_degree_list(v);
	}

	// Slide every live list to the front of the store, keeping list order.  Called when the
	// new element would not fit; (me_rest, e_rest) describe the two lists being read.
	void compact(idx &read_me, idx consumed_from_me, idx e, idx &read_e, idx left_in_e)
	{
		where[pivot] = read_me;
		length[pivot] -= consumed_from_me;
		if (length[pivot] == 0) where[pivot] = NONE;
		where[e] = read_e;
		length[e] = left_in_e;
		if (length[e] == 0) where[e] = NONE;

		for (idx v = 0; v < n; ++v) {
			const idx at = where[v];
			if (at >= 0) { where[v] = store[at]; store[at] = mirror(v); }
		}
		idx src = 0, dst = 0;
		const idx stop = pivot_first - 1;
		while (src <= stop) {
			const idx v = mirror(store[src++]);
			if (v >= 0) {
				store[dst] = where[v];
				where[v] = dst++;
				for (idx c = 0; c + 2 <= length[v]; ++c) store[dst++] = store[src++];
			}
		}
		const idx moved = dst;
		for (src = pivot_first; src <= store_free - 1; ++src) store[dst++] = store[src];
		pivot_first = moved;
		store_free = dst;
		read_e = where[e];
		read_me = where[pivot];
	}

	void form_element()
	{
		weight[pivot] = -pivot_weight;
		pivot_degree = 0;
		if (pivot_old_elems == 0) {
			// the pivot touches no element: rewrite its own variable list in place
			pivot_first = where[pivot];
			pivot_last = pivot_first - 1;
			const idx stop = pivot_first + length[pivot];
			for (idx q = pivot_first; q < stop; ++q) {
				const idx v = store[q];
				const idx w = weight[v];
				if (w > 0) { claim(v, w); store[++pivot_last] = v; }
			}
		} else {
			// merge the variable lists of the pivot's elements, then its own variables,
			// into fresh space at the end of the store
			idx read_me = where[pivot];
			pivot_first = store_free;
			const idx own_vars = length[pivot] - pivot_old_elems;
			for (idx part = 1; part <= pivot_old_elems + 1; ++part) {
				idx e, read_e, count;
				if (part > pivot_old_elems) { e = pivot; read_e = read_me; count = own_vars; }
				else { e = store[read_me++]; read_e = where[e]; count = length[e]; }
				for (idx taken = 1; taken <= count; ++taken) {
					const idx v = store[read_e++];
					const idx w = weight[v];
					if (w <= 0) continue;
					if (store_free >= store_cap) compact(read_me, part, e, read_e, count - taken);
					claim(v, w);
					store[store_free++] = v;
				}
				if (e != pivot) { where[e] = mirror(pivot); mark[e] = 0; }   // e is absorbed
			}
			pivot_last = store_free - 1;
		}
		degree[pivot] = pivot_degree;
		where[pivot] = pivot_first;
		length[pivot] = pivot_last - pivot_first + 1;
		nelem[pivot] = mirror(pivot_weight + pivot_degree);
		reset_stamp_if_needed();
	}

	// mark[e] - stamp becomes |Le \ Lme| for every element adjacent to a variable of Lme
	void external_degrees()
	{
		for (idx q = pivot_first; q <= pivot_last; ++q) {
			const idx v = store[q];
			const idx ne = nelem[v];
			if (ne <= 0) continue;
			const idx w = -weight[v];
			const idx base = stamp - w;
			const idx stop = where[v] + ne;
			for (idx t = where[v]; t < stop; ++t) {
				const idx e = store[t];
				idx m = mark[e];
				if (m >= stamp) m -= w;
				else if (m != 0) m = degree[e] + base;
				mark[e] = m;
			}
		}
	}

	void update_degrees_and_hash()
	{
		for (idx q = pivot_first; q <= pivot_last; ++q) {
			const idx v = store[q];
			const idx first = where[v];
			const idx last_elem = first + nelem[v] - 1;
			idx out = first;
			unsigned long key = 0;
			idx d = 0;
			for (idx t = first; t <= last_elem; ++t) {
				const idx e = store[t];
				const idx m = mark[e];
				if (m == 0) continue;
				const idx ext = m - stamp;
				if (!aggressive || ext > 0) {
					d += ext;
					store[out++] = e;
					key += (unsigned long)e;
				} else {
					where[e] = mirror(pivot);      // nothing of e lies outside Lme
					mark[e] = 0;
				}
			}
			nelem[v] = out - first + 1;
			const idx vars_begin = out;
			const idx stop = first + length[v];
			for (idx t = last_elem + 1; t < stop; ++t) {
				const idx u = store[t];
				const idx wu = weight[u];
				if (wu > 0) { d += wu; store[out++] = u; key += (unsigned long)u; }
			}
			if (nelem[v] == 1 && vars_begin == out) {
				// v is only connected to the pivot: eliminate it together with the pivot
				where[v] = mirror(pivot);
				const idx w = -weight[v];
				pivot_degree -= w;
				pivot_weight += w;
				eliminated += w;
				weight[v] = 0;
				nelem[v] = NONE;
			} else {
				if (d < degree[v]) degree[v] = d;
				store[out] = store[vars_begin];
				store[vars_begin] = store[first];
				store[first] = pivot;
				length[v] = out - first + 1;
				const idx bucket = (idx)(key % (unsigned long)n);
				const idx h = head[bucket];
				if (h <= NONE) { fwd[v] = mirror(h); head[bucket] = mirror(v); }
				else { fwd[v] = bwd[h]; bwd[h] = v; }
				bwd[v] = bucket;
			}
		}
		degree[pivot] = pivot_degree;
		if (pivot_degree > widest_element) widest_element = pivot_degree;
		stamp += widest_element;
		reset_stamp_if_needed();
	}

	void merge_indistinguishable()
	{
		for (idx q = pivot_first; q <= pivot_last; ++q) {
			idx v = store[q];
			if (weight[v] >= 0) continue;
			const idx bucket = bwd[v];
			const idx h = head[bucket];
			if (h == NONE) v = NONE;
			else if (h < NONE) { v = mirror(h); head[bucket] = NONE; }
			else { v = bwd[h]; bwd[h] = NONE; }

			while (v != NONE && fwd[v] != NONE) {
				const idx len_v = length[v], ne_v = nelem[v];
				for (idx t = where[v] + 1; t <= where[v] + len_v - 1; ++t) mark[store[t]] = stamp;
				idx prev = v, u = fwd[v];
				while (u != NONE) {
					bool same = (length[u] == len_v) && (nelem[u] == ne_v);
					for (idx t = where[u] + 1; same && t <= where[u] + len_v - 1; ++t)
						if (mark[store[t]] != stamp) same = false;
					if (same) {
						where[u] = mirror(v);
						weight[v] += weight[u];
						weight[u] = 0;
						nelem[u] = NONE;
						u = fwd[u];
						fwd[prev] = u;
					} else {
						prev = u;
						u = fwd[u];
					}
				}
				++stamp;
				v = fwd[v];
			}
		}
	}

	void finish_element()
	{
		idx out = pivot_first;
		const idx remaining = n - eliminated;
		for (idx q = pivot_first; q <= pivot_last; ++q) {
			const idx v = store[q];
			const idx w = -weight[v];
			if (w <= 0) continue;
			weight[v] = w;
			idx d = degree[v] + pivot_degree - w;
			if (d > remaining - w) d = remaining - w;
			push_degree_list(v, d);
			bwd[v] = NONE;
			if (d < lowest_degree) lowest_degree = d;
			degree[v] = d;
			store[out++] = v;
		}
		weight[pivot] = pivot_weight;
		length[pivot] = out - pivot_first;
		if (length[pivot] == 0) { where[pivot] = NONE; mark[pivot] = 0; }
		if (pivot_old_elems != 0) store_free = out;
	}

	void eliminate_all()
	{
		while (eliminated < n) {
			choose_pivot();
			form_element();
			external_degrees();
			update_degrees_and_hash();
			merge_indistinguishable();
			finish_element();
		}
	}

	// Assembly tree -> post-order -> permutation.
	void number_nodes(idx *perm)
	{
		std::vector<idx> &parent = where;
		std::vector<idx> &size = nelem;
		for (idx v = 0; v < n; ++v) { parent[v] = mirror(parent[v]); size[v] = mirror(size[v]); }

		// absorbed variables point straight at the element that swallowed them
		for (idx v = 0; v < n; ++v) {
			if (weight[v] != 0) continue;
			idx a = parent[v];
			if (a == NONE) continue;
			while (weight[a] == 0) a = parent[a];
			const idx e = a;
			for (idx u = v; weight[u] == 0;) { const idx up = parent[u]; parent[u] = e; u = up; }
		}

		// children lists (ascending index), biggest child moved to the back (ties: the later one)
		std::vector<idx> child(n, NONE), sibling(n, NONE), order(n, NONE), stack(n, 0);
		for (idx v = n - 1; v >= 0; --v)
			if (weight[v] > 0 && parent[v] != NONE) { sibling[v] = child[parent[v]]; child[parent[v]] = v; }
		for (idx v = 0; v < n; ++v) {
			if (weight[v] <= 0 || child[v] == NONE) continue;
			idx before = NONE, best = NONE, best_before = NONE, best_size = NONE, last = NONE;
			for (idx f = child[v]; f != NONE; f = sibling[f]) {
				if (size[f] >= best_size) { best_size = size[f]; best = f; best_before = before; }
				before = f;
				last = f;
			}
			const idx after = sibling[best];
			if (after != NONE) {
				if (best_before == NONE) child[v] = after; else sibling[best_before] = after;
				sibling[best] = NONE;
				sibling[last] = best;
			}
		}
		idx counter = 0;
		for (idx root = 0; root < n; ++root) {
			if (parent[root] != NONE || weight[root] <= 0) continue;
			idx top = 0;
			stack[0] = root;
			while (top >= 0) {
				const idx v = stack[top];
				if (child[v] != NONE) {
					for (idx f = child[v]; f != NONE; f = sibling[f]) ++top;
					idx slot = top;
					for (idx f = child[v]; f != NONE; f = sibling[f]) stack[slot--] = f;
					child[v] = NONE;
				} else {
					--top;
					order[v] = counter++;
				}
			}
		}

		// elements get consecutive blocks in post-order; an absorbed variable is put just before
		// the element it belongs to; dense variables go last in index order
		std::vector<idx> at_rank(n, NONE), position(n, NONE);
		for (idx e = 0; e < n; ++e)
			if (order[e] != NONE) at_rank[order[e]] = e;
		idx placed = 0;
		for (idx r = 0; r < n; ++r) {
			const idx e = at_rank[r];
			if (e == NONE) break;
			position[e] = placed;
			placed += weight[e];
		}
		for (idx v = 0; v < n; ++v) {
			if (weight[v] != 0) continue;
			const idx e = parent[v];
			if (e != NONE) { position[v] = position[e]; ++position[e]; }
			else position[v] = placed++;
		}
		for (idx v = 0; v < n; ++v) perm[position[v]] = v;
	}
};

} // namespace

int amd_order(long n, const long *colptr, const long *rowidx, long *perm, double dense_ratio, int aggressive)
{
	if (!colptr || !rowidx || !perm || n < 0) return AMD_STATUS_INVALID;
	if (n == 0) return AMD_STATUS_OK;
	if (colptr[n] < 0) return AMD_STATUS_INVALID;
	const int status = classify_columns(n, colptr, rowidx);
	if (status == AMD_STATUS_INVALID) return status;

	std::vector<idx> tp, ti;
	const idx *cp = colptr, *ci = rowidx;
	if (status == AMD_STATUS_JUMBLED) {
		transpose_pattern(n, colptr, rowidx, tp, ti);
		cp = tp.data();
		ci = ti.data();
	}

	QuotientGraph g;
	g.n = n;
	std::vector<idx> cursor;
	g.length.assign(n, 0);
	for_each_symmetric_edge(n, cp, ci, cursor, [&](idx a, idx b) { ++g.length[a]; ++g.length[b]; });
	std::size_t edges = 0;
	for (idx v = 0; v < n; ++v) edges += (std::size_t)g.length[v];

	// The reference sizes its workspace as nz(A+A') * 1.2 + 7n and hands all but 6n of it to
	// the adjacency store.  The capacity decides when compaction runs; compaction keeps list
	// order, so the result does not depend on it, but we keep the same capacity anyway.
	g.store_cap = (idx)(edges + edges / 5 + (std::size_t)n);
	g.store.assign(g.store_cap > 0 ? g.store_cap : 1, 0);
	g.where.assign(n, 0);
	std::vector<idx> fillpos(n, 0);
	idx at = 0;
	for (idx v = 0; v < n; ++v) { g.where[v] = at; fillpos[v] = at; at += g.length[v]; }
	g.store_free = at;
	for_each_symmetric_edge(n, cp, ci, cursor, [&](idx a, idx b) { g.store[fillpos[a]++] = b; g.store[fillpos[b]++] = a; });

	g.initialise(dense_ratio, aggressive != 0);
	g.eliminate_all();
	g.number_nodes(perm);
	return status;
}

} // namespace qpb

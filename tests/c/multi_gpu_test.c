/* A C caller of the multi-GPU entry points (include/qpswift_b200.h), linked against libqpswift_b200.so.
 *
 *     multi_gpu_test <problem.bin> <result.bin> [shards]
 *
 * problem.bin (written by tests/test_multi_gpu.py, all little-endian 64-bit): n m p count, Pjc[n+1] Pir Ajc[n+1] Air
 * Gjc[n+1] Gir, then the value rows Ppr[count][nnzP] Apr Gpr c[count][n] h[count][m] b[count][p] as doubles.
 * The program sets up ONE prototype QP the way the controller would (QP_SETUP, dogbot_controller/src/qpSWIFT/
 * qpSWIFT.c:60-234), then solves the batch (a) on device 0 alone and (b) split over `shards` shards -- one per
 * device of the box, or, with fewer devices than shards, several shards per device -- through
 * QP_SETUP_BATCH_MULTI / QP_SOLVE_BATCH_MULTI, checks that (a) and (b) agree bit for bit, and writes (b)'s
 * x y z s flag iters to result.bin for the caller to compare with the CPU oracle.  Exit code 0 = all good.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qpswift_b200.h"

static void *xmalloc(size_t bytes)
{
	void *p = malloc(bytes ? bytes : 8);
	if (!p) { fprintf(stderr, "out of memory\n"); exit(2); }
	return p;
}

static void rd(FILE *f, void *dst, size_t bytes)
{
	if (bytes && fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s problem.bin result.bin [shards]\n", argv[0]); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	qp_int hdr[4];
	rd(f, hdr, sizeof hdr);
	const qp_int n = hdr[0], m = hdr[1], p = hdr[2], count = hdr[3];
	qp_int *Pjc = xmalloc((n + 1) * sizeof(qp_int)), *Ajc = xmalloc((n + 1) * sizeof(qp_int)), *Gjc = xmalloc((n + 1) * sizeof(qp_int));
	rd(f, Pjc, (n + 1) * sizeof(qp_int));
	qp_int *Pir = xmalloc(Pjc[n] * sizeof(qp_int));
	rd(f, Pir, Pjc[n] * sizeof(qp_int));
	rd(f, Ajc, (n + 1) * sizeof(qp_int));
	qp_int *Air = xmalloc(Ajc[n] * sizeof(qp_int));
	rd(f, Air, Ajc[n] * sizeof(qp_int));
	rd(f, Gjc, (n + 1) * sizeof(qp_int));
	qp_int *Gir = xmalloc(Gjc[n] * sizeof(qp_int));
	rd(f, Gir, Gjc[n] * sizeof(qp_int));
	const size_t w[6] = {(size_t)Pjc[n], (size_t)Ajc[n], (size_t)Gjc[n], (size_t)n, (size_t)m, (size_t)p};
	qp_real *in[6];
	for (int k = 0; k < 6; ++k) {
		in[k] = xmalloc(count * w[k] * sizeof(qp_real));
		rd(f, in[k], count * w[k] * sizeof(qp_real));
	}
	fclose(f);

	const int ndev = QPB_DEVICE_COUNT();
	if (ndev < 1) { fprintf(stderr, "no CUDA device (the library has no CPU fallback)\n"); return 3; }
	int shards = argc > 3 ? atoi(argv[3]) : ndev;
	if (shards < 1) shards = 1;
	int *devices = xmalloc(shards * sizeof(int));
	for (int d = 0; d < shards; ++d) devices[d] = d % ndev;

	/* the prototype: the first QP of the batch, set up as the reference's caller does */
	QP *proto = QP_SETUP(n, m, p, Pjc, Pir, in[0], p ? Ajc : NULL, p ? Air : NULL, p ? in[1] : NULL, Gjc, Gir, in[2], in[3], in[4], p ? in[5] : NULL, 0.0, NULL);
	if (!proto || proto->stats->Flag == QP_KKTFAIL) { fprintf(stderr, "QP_SETUP failed\n"); return 4; } /* (Flag stays QP_FATAL until QP_SOLVE, qpSWIFT.c:80) */

	const size_t ow[4] = {(size_t)n, (size_t)p, (size_t)m, (size_t)m};
	qp_real *one[4], *many[4];
	for (int k = 0; k < 4; ++k) { one[k] = xmalloc(count * ow[k] * sizeof(qp_real)); many[k] = xmalloc(count * ow[k] * sizeof(qp_real)); }
	qp_int *flag1 = xmalloc(count * sizeof(qp_int)), *it1 = xmalloc(count * sizeof(qp_int));
	qp_int *flagN = xmalloc(count * sizeof(qp_int)), *itN = xmalloc(count * sizeof(qp_int));

	const int dev0 = 0;
	QPBatchMulti *A = QP_SETUP_BATCH_MULTI(proto, count, &dev0, 1);
	if (!A) { fprintf(stderr, "QP_SETUP_BATCH_MULTI(1 device): %s\n", QPB_MULTI_LAST_ERROR()); return 5; }
	qp_int bad1 = QP_SOLVE_BATCH_MULTI(A, count, proto->options, in[0], in[1], in[2], in[3], in[4], in[5], one[0], one[1], one[2], one[3], flag1, it1);
	if (bad1 < 0) { fprintf(stderr, "QP_SOLVE_BATCH_MULTI(1 device): %s\n", QPB_MULTI_LAST_ERROR()); return 6; }
	QP_CLEANUP_BATCH_MULTI(A);

	QPBatchMulti *M = QP_SETUP_BATCH_MULTI(proto, count, devices, shards);
	if (!M) { fprintf(stderr, "QP_SETUP_BATCH_MULTI(%d shards): %s\n", shards, QPB_MULTI_LAST_ERROR()); return 7; }
	qp_int badN = QP_SOLVE_BATCH_MULTI(M, count, proto->options, in[0], in[1], in[2], in[3], in[4], in[5], many[0], many[1], many[2], many[3], flagN, itN);
	if (badN < 0) { fprintf(stderr, "QP_SOLVE_BATCH_MULTI(%d shards): %s\n", shards, QPB_MULTI_LAST_ERROR()); return 8; }
	qp_int lo, hi, covered = 0;
	for (int d = 0; d < shards; ++d) {
		QPB_MULTI_SHARD(M, count, d, &lo, &hi);
		if (lo != covered || hi < lo) { fprintf(stderr, "shard %d = [%ld, %ld) is not contiguous\n", d, (long)lo, (long)hi); return 9; }
		covered = hi;
	}
	if (covered != count) { fprintf(stderr, "shards cover %ld of %ld QPs\n", (long)covered, (long)count); return 9; }
	/* count > capacity must be refused, not overrun */
	if (QP_SOLVE_BATCH_MULTI(M, count + 1, proto->options, in[0], in[1], in[2], in[3], in[4], in[5], NULL, NULL, NULL, NULL, NULL, NULL) != QPB_E_ARG) { fprintf(stderr, "count > capacity was accepted\n"); return 10; }
	QP_CLEANUP_BATCH_MULTI(M);

	int rc = 0;
	if (bad1 != badN) { fprintf(stderr, "non-optimal counts differ: %ld vs %ld\n", (long)bad1, (long)badN); rc = 11; }
	for (int k = 0; k < 4 && !rc; ++k)
		if (memcmp(one[k], many[k], count * ow[k] * sizeof(qp_real))) { fprintf(stderr, "output %d differs between 1 device and %d shards\n", k, shards); rc = 12; }
	if (!rc && (memcmp(flag1, flagN, count * sizeof(qp_int)) || memcmp(it1, itN, count * sizeof(qp_int)))) { fprintf(stderr, "flags / iteration counts differ\n"); rc = 13; }

	f = fopen(argv[2], "wb");
	if (!f) { perror(argv[2]); return 2; }
	for (int k = 0; k < 4; ++k) fwrite(many[k], sizeof(qp_real), count * ow[k], f);
	fwrite(flagN, sizeof(qp_int), count, f);
	fwrite(itN, sizeof(qp_int), count, f);
	fclose(f);
	QP_CLEANUP(proto);
	printf("multi_gpu_test: %ld QPs, %d device(s), %d shard(s), %ld not optimal, 1-device and %d-shard results %s\n", (long)count, ndev, shards, (long)badN, shards, rc ? "DIFFER" : "bit-identical");
	return rc;
}

// SURVEY 8 f4 (BASELINE.json configs[3]): the per-tick device side of a Monte-Carlo rollout -- robots x sequential
// control ticks with everything resident on the GPU: integrate the last solution, advance the contact schedule, move the
// swing feet along the attractive field towards their (perturbed) footholds, rebuild the robot's state row, and list
// the robots of each contact pattern for that pattern's formulation / solve / torque kernels (wbc.cuh, batch.cu).
//
// The reference's loop (dogbot_controller/src/client/main.cpp): ctrl_loop at 400 Hz (1107) takes the robot's state from
// Gazebo + iDynTree, the planner's CoM / swing references from towr, and the artificial-potential-field step planner
// (1261-1421: attractive field f_a = -K_pa sat(e), foothold = foot + 0.5 f_a [+ 0.5 f_r]; 2745-2845 the crawl variant)
// picks the next footholds.  None of simulator, dynamics library or trajectory optimiser exists here, so the rollout
// closes the loop with a FIXED, fully specified model (SURVEY.md 8d "a fixed integrator + APF perturbation on foothold
// targets"), every operation written out so that apf_quadruped_b200/rollout.py can step the same model in numpy and the
// oracle can solve the same QPs on the CPU, bit for bit:
//   joints      dq += DT ((sat(qdd, +-40) - 400 (q - q_mid)) - 40 dq);  q += DT dq   (semi-implicit Euler, DT = 1/400 s,
//               main.cpp:1107; the spring-damper towards mid-range stands in for the posture references towr provides --
//               the synthetic Jacobians are not a kinematic chain, so the raw qdd alone would drive the joints away)
//   CoM error   dv -= DT sat(a_com - g e_z, +-30);  dx += DT dv;  W = (3000 dx + 50 dv) + m g e_z   (main.cpp:1498-1499, 1572)
//   re-planning entering a new gait phase: dx *= 0.25, dv *= 0.25   (the planner re-anchors the CoM reference, main.cpp:1391-1421)
//   schedule    phase = ((tick + 7 robot) / 25) mod 4 -> stance, trot, stance, crawl   (towr's gait phases: fixed here)
//   footholds   entering trot / crawl: goal_f = nominal_f + U(+-0.03) for the swing feet   (the APF-perturbed targets;
//               counter-based splitmix64 of (seed, robot, tick, foot, axis): the same bits on host and device)
//   swing feet  e = sat(r_xy - goal_xy, +-0.05);  r_xy += 0.025 (0.5 (-1.5 e))          (f_a = -K_pa e, main.cpp:1331-1390)
//   tracking    vdot_sw_xy = 300 (goal - r)_xy, z: 0                                     (Kp of main.cpp:1985-1990)
//   Jacobian    Jc_f = [I | -[r_f]x] + noise_f  (noise: the robot's fixed draw)
// Everything else of the state row (Jj, M, bias, Jdqd) is the robot's fixed draw.
#pragma once

#include "wbc.cuh"

namespace qpb {

// per-robot rollout row (doubles): r[4][3] | goal[4][2] | dx[6] | dv[6] | noise[72]
enum { RR_R = 0, RR_GOAL = 12, RR_DX = 20, RR_DV = 26, RR_NOISE = 32, ROLLOUT_ROW = 104 };

struct RolloutParams
{
	double dt, kp_com, kd_com, mass_g, g, goal_spread, sat, step, half, kpa, kp_sw, qdd_sat, kq, kdq, a_sat, replan;
	int phase_ticks, robot_stride, first_robot;
	unsigned long long seed;
};

__host__ __device__ inline unsigned long long rollout_hash(unsigned long long seed, unsigned long long robot, unsigned long long tick, unsigned long long foot, unsigned long long axis)
{
	unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (robot + 1) + 0xBF58476D1CE4E5B9ull * (tick + 1) + 0x94D049BB133111EBull * (4 * foot + axis + 1);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

#ifdef __CUDACC__

__device__ __forceinline__ int rollout_kind(const long long tick, const int robot, const RolloutParams &P)
{
	const long long ph = ((tick + (long long)P.robot_stride * robot) / P.phase_ticks) & 3;
	return ph == 1 ? WBC_TROT : ph == 3 ? WBC_CRAWL : WBC_STANCE;
}

// One thread per robot.  tick = *dtick: for tick > 0 the solution x of the previous tick is integrated first.
__global__ void __launch_bounds__(128) rollout_tick_kernel(double *__restrict__ state, double *__restrict__ robot_rows, const double *__restrict__ x, int *__restrict__ kind_out, const long long *__restrict__ dtick, const int robots, const RolloutParams P)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= robots) return;
	const long long tick = *dtick;
	double *S = state + (size_t)i * WBC_STATE, *R = robot_rows + (size_t)i * ROLLOUT_ROW;
	const int gi = P.first_robot + i; // the robot's index in the whole (multi-GPU) rollout: schedule and random draws follow it
	if (tick > 0) {
		const double *xs = x + (size_t)i * 30;
		const double qmin[12] = {-1.75, -1.75, -1.75, -1.75, -1.58, -2.62, -3.15, -0.02, -1.58, -2.62, -3.15, -0.02}; // main.cpp:493
		const double qmax[12] = {1.75, 1.75, 1.75, 1.75, 3.15, 0.02, 1.58, 2.62, 3.15, 0.02, 1.58, 2.62};             // main.cpp:494
		for (int j = 0; j < 12; ++j) {
			const double qdd = fmin(fmax(xs[6 + j], -P.qdd_sat), P.qdd_sat);
			const double acc = (qdd - P.kq * (S[WS_Q + j] - 0.5 * (qmin[j] + qmax[j]))) - P.kdq * S[WS_DQ + j];
			const double dq = S[WS_DQ + j] + P.dt * acc;
			S[WS_DQ + j] = dq;
			S[WS_Q + j] = S[WS_Q + j] + P.dt * dq;
		}
		for (int k = 0; k < 6; ++k) {
			const double a = fmin(fmax(xs[k] - (k == 2 ? P.g : 0.0), -P.a_sat), P.a_sat);
			const double dv = R[RR_DV + k] - P.dt * a;
			R[RR_DV + k] = dv;
			R[RR_DX + k] = R[RR_DX + k] + P.dt * dv;
		}
	}
	const int kind = rollout_kind(tick, gi, P);
	const bool entering = tick == 0 || rollout_kind(tick - 1, gi, P) != kind;
	if (entering && tick > 0) // a new gait phase re-plans the CoM reference from where the robot is (main.cpp:1391-1421): most of the error goes
		for (int k = 0; k < 6; ++k) { R[RR_DX + k] = P.replan * R[RR_DX + k]; R[RR_DV + k] = P.replan * R[RR_DV + k]; }
	// swing feet of the kind (feet br, bl, fl, fr): trot br, fl; crawl br
	const bool swing[4] = {kind != WBC_STANCE, false, kind == WBC_TROT, false};
	const double nominal[4][2] = {{-0.186571, -0.289186}, {-0.186571, 0.289186}, {0.186571, 0.289186}, {0.186571, -0.289186}}; // main.cpp:433-434
	for (int f = 0; f < 4; ++f) {
		for (int a = 0; a < 2; ++a) {
			if (swing[f]) {
				if (entering) {
					const double u = (double)(rollout_hash(P.seed, (unsigned long long)gi, (unsigned long long)tick, f, a) >> 11) * (1.0 / 9007199254740992.0);
					R[RR_GOAL + 2 * f + a] = nominal[f][a] + (u * (2.0 * P.goal_spread) - P.goal_spread);
				}
				const double goal = R[RR_GOAL + 2 * f + a];
				const double e = fmin(fmax(R[RR_R + 3 * f + a] - goal, -P.sat), P.sat);
				const double r = R[RR_R + 3 * f + a] + P.step * (P.half * (-P.kpa * e));
				R[RR_R + 3 * f + a] = r;
				S[WS_VSW + 3 * f + a] = P.kp_sw * (goal - r);
			} else S[WS_VSW + 3 * f + a] = 0.0;
		}
		S[WS_VSW + 3 * f + 2] = 0.0;
		// Jc_f = [I | -[r_f]x] + noise_f
		const double rx = R[RR_R + 3 * f], ry = R[RR_R + 3 * f + 1], rz = R[RR_R + 3 * f + 2];
		const double base[3][6] = {{1.0, 0.0, 0.0, 0.0, rz, -ry}, {0.0, 1.0, 0.0, -rz, 0.0, rx}, {0.0, 0.0, 1.0, ry, -rx, 0.0}};
		for (int a = 0; a < 3; ++a)
			for (int c = 0; c < 6; ++c) S[WS_JC + 6 * (3 * f + a) + c] = base[a][c] + R[RR_NOISE + 6 * (3 * f + a) + c];
	}
	for (int k = 0; k < 6; ++k) S[WS_W + k] = (P.kp_com * R[RR_DX + k] + P.kd_com * R[RR_DV + k]) + (k == 2 ? P.mass_g : 0.0);
	kind_out[i] = kind;
}

// The robots of each contact pattern, in ascending order (one CTA: an ordered compaction, so that the lists -- and with
// them the order in which warps take QPs -- do not depend on scheduling), their counts, and the tick counter's step.
__global__ void __launch_bounds__(1024) rollout_lists_kernel(const int *__restrict__ kind, int *__restrict__ lists, int *__restrict__ counts, long long *__restrict__ dtick, const int robots)
{
	__shared__ int warp_tot[3][32];
	__shared__ int base[3];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
	if (threadIdx.x < 3) base[threadIdx.x] = 0;
	__syncthreads();
	for (int i0 = 0; i0 < robots; i0 += blockDim.x) {
		const int i = i0 + threadIdx.x;
		const int k = i < robots ? kind[i] : -1;
		unsigned before[3];
		for (int c = 0; c < 3; ++c) {
			const unsigned m = __ballot_sync(0xffffffffu, k == c);
			before[c] = __popc(m & ((1u << lane) - 1u));
			if (lane == 0) warp_tot[c][warp] = __popc(m);
		}
		__syncthreads();
		if (k >= 0) {
			int off = base[k];
			for (int w = 0; w < warp; ++w) off += warp_tot[k][w];
			lists[(size_t)k * robots + off + before[k]] = i;
		}
		__syncthreads();
		if (threadIdx.x < 3) {
			int t = 0;
			for (int w = 0; w < warps; ++w) t += warp_tot[threadIdx.x][w];
			base[threadIdx.x] += t;
		}
		__syncthreads();
	}
	if (threadIdx.x < 3) counts[threadIdx.x] = base[threadIdx.x];
	if (threadIdx.x == 0) *dtick = *dtick + 1;
}

#endif // __CUDACC__

} // namespace qpb

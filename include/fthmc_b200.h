/* fthmc_b200.h -- C ABI of libfthmc_b200.so: the B200 (sm_100a) implementation of nftqcd/fthmc's
 * field-transformed HMC trajectory path for 2D U(1) lattice gauge theory.
 *
 * The reference (pure Python on PyTorch) has no FFI layer; its boundary for this path is the set of
 * Python functions cited below (paths relative to the reference root).  Each export here is what a
 * binding for that function calls; INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns every buffer,
 *     including the workspace (size it with fthmc_workspace_bytes); the library allocates device
 *     memory only inside fthmc_flow_pack (the packed weights) and frees it in fthmc_flow_free;
 *   - link fields are contiguous row-major (B,2,L0,L1), the reference's layout; per-chain scalars are (B);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - return value: 0 ok, >0 a cudaError_t, <0 an argument error (FTHMC_E_*); the message is kept in
 *     fthmc_last_error_string() (thread-local).  There is NO CPU fallback and no silent dispatch.
 *   - L0 and L1 must be multiples of 4 (the 4-periodic stripe masks, ipynb/field_transformation.py:175-248);
 *     the resident-chain entry points keep one chain in one SM's shared memory (L0*L1 <= 1024 sites in fp64) or, for
 *     larger lattices (L = 48 .. 128), in the distributed shared memory of one thread-block cluster of up to 16 CTAs;
 *     beyond that they return FTHMC_E_LATTICE.
 */
#ifndef FTHMC_B200_H
#define FTHMC_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTHMC_OK            0
#define FTHMC_E_ARG        (-1)   /* null pointer / non-positive size / unknown flag */
#define FTHMC_E_LATTICE    (-2)   /* L % 4 != 0, or lattice too large for the shared-memory-resident path */
#define FTHMC_E_DTYPE      (-3)   /* unsupported dtype for this entry point */
#define FTHMC_E_WORKSPACE  (-4)   /* workspace pointer null or too small */
#define FTHMC_E_NETSHAPE   (-5)   /* CNN shape other than 2 -> 8 -> 8 -> (2+1), kernel 3 */

enum { FTHMC_F64 = 0, FTHMC_F32 = 1 };
enum { FTHMC_ACT_SILU = 0, FTHMC_ACT_LEAKY_RELU = 1, FTHMC_ACT_RELU = 2 };    /* fthmc/utils/layers.py:117-122 */
enum { FTHMC_CONV_0_2PI = 0,      /* torch_mod in [0,2pi):  ipynb/field_transformation.py:17-18 (scripts) */
       FTHMC_CONV_MPI_PI = 1 };   /* torch_mod in [-pi,pi): fthmc/utils/layers.py:41-43 (package)         */
enum { FTHMC_ORDER_U1PLAQ = 0,    /* compute_u1_plaq term order, ipynb/field_transformation.py:118-119    */
       FTHMC_ORDER_PLAQPHASE = 1 };/* plaqphase term order, hmc_2dU1.py:114-120                            */

typedef struct fthmc_flow* fthmc_flow_t;   /* immutable packed flow (device resident) */

int          fthmc_version(void);
const char*  fthmc_last_error_string(void);

/* ---- plain Wilson stencils (HBM-bound streaming kernels) --------------------------------------------- */
/* action(param,f) hmc_2dU1.py:100 / U1GaugeAction ipynb/field_transformation.py:120: out[b] = -beta*sum cos P */
int fthmc_action(const void* links, int B, int L0, int L1, double beta, int order, void* out, int dtype, void* stream);
/* force(param,f) hmc_2dU1.py:104 (closed form of the autograd result): force_out (B,2,L0,L1) */
int fthmc_force(const void* links, int B, int L0, int L1, double beta, int order, void* force_out, int dtype, void* stream);
/* rounded=1: topocharge(f) hmc_2dU1.py:123 (floor(0.1+sum regularize(P)/2pi));
 * rounded=0: topo_charge(x) ipynb/field_transformation.py:138 (un-rounded sum of wrapped plaquettes / 2pi) */
int fthmc_topo_charge(const void* links, int B, int L0, int L1, int rounded, void* out, int dtype, void* stream);
/* regularize(f) hmc_2dU1.py:127: elementwise wrap to [-pi,pi); in place allowed */
int fthmc_regularize(const void* in, void* out, long long n, int dtype, void* stream);

/* ---- workspace --------------------------------------------------------------------------------------- */
/* bytes needed by any resident-chain call below for B chains (flow may be NULL for the plain-HMC calls) */
size_t fthmc_workspace_bytes(fthmc_flow_t flow, int B, int L0, int L1);

/* ---- plain HMC, one CTA per chain, lattice resident in shared memory (fp64) ---------------------------- */
/* leapfrog(param,x,p) hmc_2dU1.py:132 */
int fthmc_leapfrog(const double* x_in, const double* p_in, double* x_out, double* p_out, int B, int L0, int L1,
                   double beta, double dt, int nstep, void* ws, size_t ws_bytes, void* stream);
/* hmc(param,x) hmc_2dU1.py:144.  p_in/u_in NULL => momenta/uniforms from Philox(seed, chain0+b, traj).
 * Outputs (any may be NULL): dH, exp(-dH), acc (0/1), plaq of the returned field, floored topological charge. */
int fthmc_hmc_traj(const double* x_in, double* x_out, const double* p_in, const double* u_in,
                   unsigned long long seed, unsigned long long traj, unsigned long long chain0,
                   int B, int L0, int L1, double beta, double dt, int nstep,
                   double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- flow -------------------------------------------------------------------------------------------- */
/* raw_host: n_layers x 955 doubles in the reference's parameter order of layer.plaq_coupling.net
 * (conv0.weight (8,2,3,3), conv0.bias, conv1.weight (8,8,3,3), conv1.bias, conv2.weight (3,8,3,3), conv2.bias;
 * make_conv_net ipynb/field_transformation.py:84-99); mu/off: the mask parameters of each layer
 * (make_u1_equiv_layers :343-344).  Synchronous (it uploads). */
int fthmc_flow_pack(const double* raw_host, int n_layers, const int* mu_host, const int* off_host,
                    int hidden0, int hidden1, int n_mix, int ksize, int activation, int convention,
                    double inv_tol, int inv_max_iter, fthmc_flow_t* out);
/* new weights (same n_layers / masks) into an existing handle; synchronous, no launch on the handle may be in flight */
int fthmc_flow_update(fthmc_flow_t flow, const double* raw_host);
int fthmc_flow_free(fthmc_flow_t flow);
int fthmc_flow_n_layers(fthmc_flow_t flow);

/* ft_flow(flow,f) ipynb/ft_hmc.py:220; logJ (B) and layer_logJ (B,n_layers) optional (sum / per-layer logJ
 * of GaugeEquivCouplingLayer.forward, ipynb/field_transformation.py:160) */
int fthmc_flow_fwd(fthmc_flow_t flow, const double* x_in, double* x_out, double* logJ, double* layer_logJ,
                   int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
/* ft_flow_inv(flow,f) ipynb/ft_hmc.py:225 (per-chain bisection, invert_transform_bisect :263);
 * iters (B,n_layers) optional: bisection iterations used */
int fthmc_flow_inv(fthmc_flow_t flow, const double* x_in, double* x_out, double* logJ, double* layer_logJ, int* iters,
                   int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
/* ft_action(param,flow,f) ipynb/ft_hmc.py:230: out[b] = S(F(x)) - sum logJ; flowed (B,2,L0,L1) optional = F(x) */
int fthmc_ft_action(fthmc_flow_t flow, const double* x, double beta, double* out, double* flowed,
                    int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
/* ft_force(param,flow,field) ipynb/ft_hmc.py:240: hand-written adjoint instead of torch.autograd */
int fthmc_ft_force(fthmc_flow_t flow, const double* x, double beta, double* force_out,
                   int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
/* ft_leapfrog(param,flow,x,p) ipynb/ft_hmc.py:394 (without its per-step diagnostics) */
int fthmc_ft_leapfrog(fthmc_flow_t flow, const double* x_in, const double* p_in, double* x_out, double* p_out,
                      int B, int L0, int L1, double beta, double dt, int nstep, void* ws, size_t ws_bytes, void* stream);
/* ft_hmc(param,flow,field) ipynb/ft_hmc.py:420, one persistent CTA per chain.  p_in/u_in as in fthmc_hmc_traj.
 * h0/h1 optional: the two Hamiltonians (diagnostics). */
int fthmc_ft_hmc_traj(fthmc_flow_t flow, const double* field_in, double* field_out, const double* p_in, const double* u_in,
                      unsigned long long seed, unsigned long long traj, unsigned long long chain0,
                      int B, int L0, int L1, double beta, double dt, int nstep,
                      double* dH, double* exp_mdH, int* acc, double* plaq, double* topo, double* h0, double* h1,
                      void* ws, size_t ws_bytes, void* stream);

/* ---- run loops -------------------------------------------------------------------------------------------- */
/* The trajectory loops of run(param, field) hmc_2dU1.py:697-707 / ipynb/ft_hmc.py:199-208 and ft_run(param, flow, field)
 * ipynb/ft_hmc.py:454-467: ntraj consecutive trajectories of every chain in ONE launch (one per device wave of chains), the field resident in shared
 * memory throughout.  Per-trajectory arrays are (ntraj, B) -- dH, exp(-dH), acc, and the two observables the
 * reference recomputes after every trajectory (plaq = action/(-beta V), floored topological charge) -- and p_in / u_in,
 * when given, are (ntraj, B, 2, L0, L1) / (ntraj, B) in trajectory order; NULL => Philox(seed, chain0+b, traj0+t). */
int fthmc_hmc_run(const double* x_in, double* x_out, const double* p_in, const double* u_in,
                  unsigned long long seed, unsigned long long traj0, unsigned long long chain0,
                  int B, int L0, int L1, double beta, double dt, int nstep, int ntraj,
                  double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                  void* ws, size_t ws_bytes, void* stream);
int fthmc_ft_hmc_run(fthmc_flow_t flow, const double* field_in, double* field_out, const double* p_in, const double* u_in,
                     unsigned long long seed, unsigned long long traj0, unsigned long long chain0,
                     int B, int L0, int L1, double beta, double dt, int nstep, int ntraj,
                     double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- flow training gradient ---------------------------------------------------------------------------- */
/* The reverse-KL training step (train_step, ipynb/ft_hmc.py:253-295; fthmc/train.py:162-228) minimises
 * mean_b [logq - logp] = mean_b ft_action(xi_b) + const over prior samples xi_b, and gets d/dweights from loss.backward().
 * fthmc_ft_action_grad computes action_out[b] = ft_action(x_b) and the gradient of sum_b ft_action(x_b) with respect to
 * the CNN weights of every layer in ONE launch: the input-gradient sweep of ft_force extended by the weight-gradient
 * GEMMs (fp64 tensor path).  grad_canon (device, n_layers x fthmc_grad_doubles()) is in the kernels' packed layout;
 * fthmc_grad_unpack (host) converts it to the reference's parameter order (n_layers x 955, the layout of raw_host in
 * fthmc_flow_pack).  force_out (B,2,L0,L1) optional: d/dx of the same sum.  Needs L0, L1 multiples of 8, L0*L1 <= 1024. */
size_t fthmc_grad_workspace_bytes(fthmc_flow_t flow, int B, int L0, int L1);
int fthmc_ft_action_grad(fthmc_flow_t flow, const double* x, double beta, double* action_out, double* grad_canon, double* force_out,
                         int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
/* Vector-Jacobian product of the flow map (x, weights) -> (y = F(x), logJ), i.e. what loss.backward() propagates through
 * `layer.forward` of every coupling layer (ipynb/field_transformation.py:160-166, 300-317) in the reference's train_step
 * (ipynb/ft_hmc.py:253-295) for ANY loss built on (y, logJ): grad_canon / grad_x receive d/dweights (canonical layout, see
 * fthmc_grad_unpack) and d/dx of  sum_b [ <gy_b, y_b> + glj_b * logJ_b ].  gy (B,2,L0,L1), glj (B); grad_x may be null.
 * fthmc_ft_action_grad is the special case gy = dS/dy, glj = -1.  Workspace: fthmc_grad_workspace_bytes. */
int fthmc_flow_vjp(fthmc_flow_t flow, const double* x, const double* gy, const double* glj, double* grad_canon, double* grad_x,
                   int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream);
int fthmc_grad_doubles(void);
int fthmc_grad_unpack(const double* grad_canon_host, int n_layers, const int* mu_host, double* raw_host);

/* diagnostic: launch `blocks` x 256 threads of pure fp64 FMA chains (2*16*iters flop per thread); *flop_out (host)
 * receives the flop count.  Used by bench.py to measure the fp64 roofline denominator on the device. */
int fthmc_diag_dfma_probe(void* scratch, int iters, int blocks, void* stream, double* flop_out_host);
/* the same on the fp64 tensor path: 8 independent DMMA.8x8x4 accumulator tiles per warp (512 flop per DMMA per warp) */
int fthmc_diag_dmma_probe(void* scratch, int iters, int blocks, void* stream, double* flop_out_host);

/* diagnostic: CTAs a chain of this lattice is spread over on the current device -- 1 = the whole chain resident in one SM's
 * shared memory (L0*L1 <= 1024 with a flow), 2..16 = one thread-block cluster per chain, 0 = does not fit / no device */
int fthmc_chain_ranks(int L0, int L1, int with_flow);

/* number of kernel launches this library has issued in this process (bench.py's gpu_launches) */
unsigned long long fthmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif

/* visnet_b200.h -- C ABI of the B200-native ViSNet energy/force engine.
 *
 * Drop-in boundary for the one hot path of microsoft/AI2BMD: the per-MD-step ViSNet evaluation over a
 * packed batch of protein fragments.  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only -- no torch types.  Every
 * function returns 0 on success or a negative vb_status; the message is available from vb_last_error().
 * There is no CPU fallback: every compute entry fails with VB_ERR_CUDA when no sm_100 device is usable.
 *
 * Units/dtypes are the reference's: positions in Angstrom, energies in eV, forces in eV/Angstrom, fp32.
 */
#ifndef VISNET_B200_H
#define VISNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vb_handle vb_handle;

typedef enum {
    VB_OK = 0,
    VB_ERR_ARG = -1,      /* bad argument / hyper-parameter mismatch */
    VB_ERR_CUDA = -2,     /* CUDA runtime error (no device, launch failure, ...) */
    VB_ERR_STATE = -3,    /* call order (e.g. forward before set_topology) */
    VB_ERR_ALLOC = -4
} vb_status;

/* Hyper-parameters of the checkpoint (src/ViSNet/model/visnet.py:14-30; both shipped checkpoints:
 * embedding_dimension 128, num_layers 6, num_heads 8, num_rbf 32, lmax 1, cutoff 5.0,
 * max_num_neighbors 32).  The kernels are specialised for exactly these; vb_create() rejects others. */
typedef struct {
    int32_t hidden_channels;
    int32_t num_layers;
    int32_t num_heads;
    int32_t num_rbf;
    int32_t max_num_neighbors;
    float cutoff;
} vb_hparams;

/* Order and element counts ("name:count;...") of the flat fp32 weight blob vb_create() expects.
 * Replaces: ViSNet.load_state_dict in load_model(), src/ViSNet/model/visnet.py:73-93. */
const char* vb_weight_manifest(void);

/* Create an engine on CUDA device `device` from a host weight blob laid out per vb_weight_manifest().
 * Replaces: get_visnet_model(model_path, device) / ViSNetModel.__init__,
 *           src/Calculators/visnet_calculator.py:36-45,184-204. */
int vb_create(const float* weights_host, size_t n_floats, const vb_hparams* hp, int device, vb_handle** out);
void vb_destroy(vb_handle* h);
const char* vb_last_error(const vb_handle* h);   /* h may be NULL: last creation error */

/* Static topology of the packed batch: atomic numbers and graph ids (sorted, contiguous) of N atoms in
 * G fragments -- host pointers, copied.  max_edges <= 0 selects the worst case N*32.
 * Replaces: the z / batch members of FragmentData (src/AIMD/fragment.py:7-13) that
 *           ViSNetModel.collate() uploads every step (visnet_calculator.py:47-52). */
int vb_set_topology(vb_handle* h, int64_t n_atoms, int64_t n_graphs, const int64_t* z_host,
                    const int64_t* batch_host, int64_t max_edges);

/* One evaluation, device buffers, asynchronous on `stream` (a cudaStream_t passed as void*).
 *   pos_dev[N*3] -> energy_dev[G], forces_dev[N*3].
 * Replaces: ViSNet.forward, src/ViSNet/model/visnet.py:135-166 (energy + autograd force). */
int vb_forward(vb_handle* h, const float* pos_dev, float* energy_dev, float* forces_dev, void* stream);

/* One evaluation with HOST buffers (pinned staging + H2D/D2H inside), synchronous.
 * Replaces: ViSNetModel.dl_potential_loader(FragmentData) -> (e[G,1], f[N,3]),
 *           src/Calculators/visnet_calculator.py:54-63. */
int vb_forward_host(vb_handle* h, const float* pos_host, float* energy_host, float* forces_host);

/* Whole-protein reduction map: F_prot[dst_atom[m]] += sign[m] * F[src_atom[m]], E_prot = sum_g frag_sign[g]*E_g.
 * Replaces: DipeptideBondedCombiner.energy_combine / forces_combine, src/Calculators/combiner.py:11-41
 *           (select_index / origin_index built at src/Fragmentation/distancefrag.py:335-353) and the
 *           dipeptide / ACE-NME split of src/Calculators/bonded.py:91-93. */
int vb_set_protein_map(vb_handle* h, int64_t n_protein_atoms, int64_t n_map, const int32_t* src_atom_host,
                       const int32_t* dst_atom_host, const float* sign_host, const float* frag_sign_host);

/* Evaluation + signed scatter into ef_prot_dev[3*n_protein_atoms + 1] (forces, then the energy in the last
 * slot); the buffer is overwritten.  With several GPUs each rank calls this on its shard of fragments and
 * the caller all-reduces ef_prot_dev (NCCL sum).  Replaces: DLBondedCalculator.__call__, bonded.py:102-123. */
int vb_forward_protein(vb_handle* h, const float* pos_dev, float* ef_prot_dev, void* stream);

/* Copy the current neighbour list to the host: slots[N*32] (source index or -1), deg[N].
 * Replaces: the edge_index returned by torch_cluster.radius_graph at src/ViSNet/model/utils.py:260-266. */
int vb_get_edges(vb_handle* h, int32_t* slots_host, int32_t* deg_host);

/* Number of kernel launches of one vb_forward(), and whether it replays a captured CUDA graph. */
int vb_launches_per_forward(const vb_handle* h);
/* Tuning knobs: "use_graph" 0/1, "npw" 1/2, "te_fwd" 32/64, "te_bwd" 32/64, "node_impl" 0/1,
 * "edge_tc" bit0 = forward / bit1 = adjoint edge stage on tcgen05 (default: chosen by problem size). */
int vb_set_option(vb_handle* h, const char* key, int64_t value);
int64_t vb_get_option(const vb_handle* h, const char* key);   /* resolved value (after vb_set_topology) */

/* ---- diagnostics (stage-by-stage parity checks; not part of the hot path) ---- */
int vb_num_stages(const vb_handle* h);
const char* vb_stage_name(const vb_handle* h, int stage);
/* Run only the first n_stages launches of an evaluation, synchronously (no graph). */
int vb_debug_run(vb_handle* h, const float* pos_dev, int n_stages);
/* Per-launch device time (ms, CUDA events on the launching stream, average of n_iter eager evaluations after
 * one warm-up) for each of the vb_num_stages() launches; used by bench.py for the live roofline numbers. */
int vb_profile_stages(vb_handle* h, const float* pos_dev, int n_iter, float* ms_per_stage_host);
/* Self-test of the tcgen05/TMEM/TMA GEMM pipeline: d[128][128] = a[128][128] * W^T, W given as a tensor-core
 * weight image (ai2bmd_b200.weights.tc_image); repeated `reps` times inside one launch; *ms_out = kernel time. */
int vb_tc_selftest(int device, const float* a_host, const float* img_host, float* d_host, int reps, float* ms_out);
/* Copy an internal buffer to the host.  name: "X","V","F","VN","QKV","V123","VDOT","TU","O" (per layer),
 * "XA","VA","GX","GVEC","GF","GXA","GQKV","GVNMSG","GTU","geom","rbf","eacc","grbf","esrc","edst","rowptr",
 * "eatom","energy","forces".  Returns the number of bytes copied (<= cap_bytes) or a negative status. */
int64_t vb_debug_read(vb_handle* h, const char* name, int layer, void* host_dst, int64_t cap_bytes);

#ifdef __cplusplus
}
#endif
#endif /* VISNET_B200_H */

/* gcpb200.h -- C ABI of the B200-native GCP-tree CEM rollout library (libgcpb200.so).
 *
 * This is the drop-in boundary: each entry point replaces one piece of the reference's Python hot path
 * (paths relative to the reference repository orybkin/video-gcp):
 *
 *   gcpb200_seq_rollout    SequentialModel forward in val_mode: SequentialRecModule.forward + VRNNCell prior branch +
 *                          decode_seq + run_auxilliary_models (gcp/prediction/models/sequential.py:33-58;
 *                          blox/torch/models/vrnn.py:54-110; blox/torch/recurrent_modules.py:21-53,195-259)
 *   gcpb200_rollout        BaseGCPModel.forward in val_mode, i.e. run_encoder + get_end_ind +
 *                          TreeModel.predict_sequence + run_auxilliary_models
 *                          (gcp/prediction/models/base_gcp.py:140-161,184-262,
 *                           gcp/prediction/models/tree/tree.py:42-67)
 *   gcpb200_prune_gather   BalancedEvalBinding.get_all_samples / __call__ + pad_sequence
 *                          (gcp/evaluation/evaluation_matching.py:174-206; base_gcp.py:361-374,242)
 *   gcpb200_cost_l2        L2ImageCost._compute + CostFcn.__call__ (gcp/planning/cem/cost_fcn.py:9-22,65-72)
 *   gcpb200_cost_learned   ImageWrappedLearnedCostFcn / LearnedCostEstimate list branch
 *                          (gcp/planning/cem/cost_fcn.py:79-116) with TestTimeCostModel.forward
 *                          (gcp/prediction/models/auxilliary_models/cost_mdl.py:138-145)
 *   gcpb200_cost_pairs     LearnedCostEstimate ndarray / list branches on row pairs, as HierarchicalTreeLatentOptimizer uses them
 *                          (gcp/planning/cem/cost_fcn.py:84-97; gcp/planning/tree_optimizer.py:96-99,145-150)
 *   gcpb200_infer_action   ImageCEMPolicy._infer_action: encoder + inverse model (gcp/planning/planner_policy.py:215-221)
 *   gcpb200_forward_loss   training-phase forward + loss: BaseGCPModel.forward(phase='train') with the approximate posterior,
 *                          TreeModel.loss and get_total_loss, as train.py:155-157 / :204-206 call them
 *                          (gcp/prediction/models/base_gcp.py:140-304; tree/tree.py:42-79; tree/tree_module.py:67-157;
 *                           tree/inference.py:16-41; tree/frame_binding.py:42-100)
 *   gcpb200_cdist_mean / gcpb200_soft_dtw / gcpb200_dtw / gcpb200_gather_rows
 *                          the DTW family: AdaptiveBinding.get_w and DTWEvalBinding.get_single_matches (declared at the end)
 *   gcpb200_topk           CEMPlanner._get_best_rollouts argsort + slice (gcp/planning/cem/cem_planner.py:124-135)
 *   gcpb200_refit          FlatCEMSampler.fit (gcp/planning/cem/sampler.py:44-46)
 *   gcpb200_sample_noise   FlatCEMSampler.sample (gcp/planning/cem/sampler.py:40-42), on device
 *
 * Conventions: plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name says
 * host; the caller owns all I/O buffers, the context owns packed weights and workspace; every call is
 * stream-ordered on `stream` (a cudaStream_t passed as void*) and never synchronises the device; return
 * value 0 = success, non-zero = error with the message available from gcpb200_last_error(); no C++
 * exception crosses this boundary.  One context per (device, host thread).  gcpb200_rollout forks work onto streams
 * the context owns (noise upload, per-call decoder constants, projection GEMMs of small tree levels) and joins every
 * one of them back into `stream` by events before the results that depend on them: to the caller the call is ordered
 * on `stream` alone.
 */
#ifndef GCPB200_H
#define GCPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcpb200_ctx gcpb200_ctx;

typedef struct {
    int device;          /* CUDA device ordinal */
    int max_candidates;  /* largest B of one rollout call (workspace is sized for it, rounded up to 128) */
    int attach_cost_mdl; /* 1: expect cost_mdl.cost_pred.* weights (learned cost available) */
    int reserved0;       /* must be 0 (the separately built verification library of tests/cuda uses it to select its
                            SIMT cross-check kernels; the shipped library has one code path and rejects non-zero) */
    int decoder_slot_chunk; /* decoder processes this many tree slots per pass (0 = default 64) */
    int model;           /* GCPB200_MODEL_TREE (0), _SEQUENTIAL (1) or _TREE_ADAPTIVE (2): which reference model the context
                            holds (TreeModel, gcp/prediction/models/tree/tree.py:14; SequentialModel,
                            gcp/prediction/models/sequential.py:104) */
    /* Tree shape; 0 = the 25-room values.  The 9-room planner config (experiments/control/9room/gcp_tree/mod_hyper.py:33-54)
     * is hierarchy_levels 7, max_seq_len 100, tied_layers 1. */
    int hierarchy_levels; /* tree depth d in [2,8]: 2^d - 1 nodes per candidate (default 8) */
    int max_seq_len;      /* frames per sequence, a multiple of 4, <= 2^d - 1 for the tree models (default 200); every
                             [B,200,..] / [B,255,..] extent in this header scales with these two */
    int tied_layers;      /* 1: one TreeModule for all levels (untied_layers False: state-dict keys tree_module.*, not
                             tree_module.tree_modules.<level>.*; gcp/prediction/models/tree/tree.py:18-21) */
} gcpb200_config;

#define GCPB200_MODEL_TREE 0
#define GCPB200_MODEL_SEQUENTIAL 1
#define GCPB200_MODEL_TREE_ADAPTIVE 2   /* TreeModel with base_configs/gcp_adaptive.py: pixel-copy decoder
                                           (blox/torch/encoder_decoder.py:235-259) and AdaptiveBinding pruning
                                           (gcp/prediction/models/adaptive_binding/adaptive.py:62-77) */
#define GCPB200_SEQ_STEPS 199   /* max_seq_len - 1 predicted frames (sequential.py:50) */

/* one fp32 host tensor of the reference state dict */
typedef struct {
    const char* name;   /* reference state-dict key */
    const float* data;  /* HOST pointer, contiguous fp32 */
    int ndim;
    int64_t shape[4];
} gcpb200_tensor;

typedef struct {
    /* ---- inputs ---- */
    const float* I_0;        /* [B,3,32,32] (or [1,3,32,32] if images_shared) fp32 in [-1,1] */
    const float* I_g;
    int images_shared;       /* 1: every candidate has the same start/goal image (a CEM call) */
    const float* z;          /* [B,255,256] noise, depth-first node order (device) */
    const float* z_host;     /* optional PINNED HOST copy of the noise: if non-NULL, `z` is only the device staging
                                buffer; the library uploads z_host -> z level by level on its own copy stream,
                                overlapped with the encoder, the tree and the decoder of the finished tree levels
                                (the decoder runs level-ordered: levels 0-5, then 6, then 7), so a host-noise call
                                takes the same time as a device-noise call (cem_simulator.py:19-26 does this copy up
                                front with torch.tensor(..., device)) */
    const int64_t* end_ind;  /* [B] injected rollout length, or NULL: sample from the length predictor */
    uint64_t seed;           /* RNG seed for length sampling */
    int B;
    /* ---- outputs (any may be NULL) ---- */
    float* e_0;              /* [B,128] */
    float* e_g;              /* [B,128] */
    float* seq_len_logits;   /* [B,200] */
    int64_t* end_ind_out;    /* [B] */
    float* e_df;             /* [B,255,128] node latents, depth-first (tree.df.e_g_prime) */
    float* mu_df;            /* [B,255,256] prior mean */
    float* log_sigma_df;     /* [B,255,256] prior log sigma */
    float* images_df;        /* [B,255,3,32,32] decoded node images (tree.df.images) */
    float* existence;        /* [B,255] existence-predictor logits */
    float* model_enc_seq;    /* [B,200,128] pruned latents, zero padded */
    float* actions;          /* [B,200,2] inverse model on consecutive pruned latents (use [:, :Lmax-1]) */
    float* regressed_state;  /* [B,200,2] state regressor (use [:, :Lmax]) */
    /* ---- GCPB200_MODEL_TREE_ADAPTIVE only (existence / model_enc_seq / actions / regressed_state must be NULL) ---- */
    float* distances;        /* [B,254] distance-predictor logits of consecutive depth-first nodes */
    int32_t* pruned_nodes;   /* [B,255] depth-first indices of the kept nodes, in order (first pruned_len[c] valid) */
    int32_t* pruned_len;     /* [B] number of kept nodes: node 0 and every node n with sigmoid(distance[n-1]) <= threshold */
    float prune_threshold;   /* learned_pruning_threshold (0 -> the reference default 0.5) */
    /* ---- planner mode, GCPB200_MODEL_TREE only (all zero / NULL = the reference behaviour: every node decoded) ----
     * The CEM planner reads only the frames balanced pruning keeps (end_ind + 1 of the 255 nodes, known before the decoder
     * runs; cem_simulator.py:29-61, tree.py:52-65) and, with the L2 image cost, only their distance to the goal image
     * (cost_fcn.py:65-72). */
    int decode_kept_only;    /* 1: decode only the kept nodes.  images_df (if given) receives exactly those nodes' images at
                                their depth-first positions -- the entries of pruned-away nodes are left untouched -- so
                                gcpb200_prune_gather / gcpb200_cost_l2 on it give the same bits as after a full decode */
    const float* l2_goal;    /* [3,32,32] goal image in [-1,1] or NULL: with l2_cost, the L2 image cost is reduced inside
                                the decoder-tail kernel from the accumulators (images_df may then be NULL: no image is
                                written at all) */
    float* l2_cost;          /* [B] L2ImageCost of every candidate (cost_fcn.py:9-22,65-72); same bits in both decode modes */
    int l2_dense;            /* dense_cost: sum over frames (1) or last frame only (0) */
    float l2_final_step_weight;
    int tree_kept_only;      /* 1 (needs decode_kept_only, no existence output): the TREE recursion, too, visits only the (node,
                                128-candidate tile) pairs some candidate keeps -- level 7 is half of the recursion and a tenth
                                of its nodes is kept on average.  e_df / mu_df / log_sigma_df then hold valid entries for
                                kept nodes only; costs, pruned frames and the pruned-sequence heads are bit-identical.  Most
                                effective when neighbouring candidates have similar lengths */
    int sort_sampled_lengths; /* 1 (images_shared, end_ind == NULL): the sampled rollout lengths are handed to the candidates in
                                descending order.  Every candidate's length is an i.i.d. draw from the same predictor output and
                                independent of its noise, so the joint law of (noise, length) over the batch -- hence costs,
                                elites and refit of a CEM iteration -- is unchanged; candidate tiles become homogeneous in
                                length, which is what tree_kept_only needs to skip work */
} gcpb200_rollout_io;

/* I/O of the sequential GCP rollout (SequentialModel forward in val_mode with injected z, default phase as the
 * simulator calls it: gcp/prediction/models/sequential.py:33-58,78-94; blox/torch/models/vrnn.py:54-110;
 * base_gcp.py:140-161,234-262).  The inference LSTM / q(z) of the VRNN cell do not reach any rollout output
 * and are not computed. */
typedef struct {
    /* ---- inputs ---- */
    const float* I_0;        /* [B,3,32,32] (or [1,3,32,32] if images_shared) fp32 in [-1,1] */
    const float* I_g;
    int images_shared;
    const float* z;          /* [B,199,256] noise, one row per predicted frame (device) */
    const int64_t* end_ind;  /* [B] injected predicted rollout length, or NULL: sample from the length predictor */
    const int64_t* given_end_ind; /* [B] inputs.end_ind (the simulator passes rollout_len-1), or NULL = 199: length of
                                     cat(e_0, encodings) kept in model_enc_seq (get_matched_pruned_seqs, base_gcp.py:361-374) */
    uint64_t seed;
    int B;
    /* ---- outputs (any may be NULL) ---- */
    float* e_0;              /* [B,128] */
    float* e_g;              /* [B,128] */
    float* seq_len_logits;   /* [B,200] */
    int64_t* end_ind_out;    /* [B] */
    float* encodings;        /* [B,199,128] predicted latents (outputs.dense_rec.encodings) */
    float* mu;               /* [B,199,256] prior mean per step */
    float* log_sigma;        /* [B,199,256] */
    float* images;           /* [B,200,3,32,32]: frame 0 = I_0, frames 1..199 decoded (outputs.dense_rec.images) */
    float* model_enc_seq;    /* [B,200,128] cat(e_0, encodings)[:given_end_ind+1], zero padded */
    float* actions;          /* [B,200,2] inverse model on consecutive rows of model_enc_seq (use [:, :Lmax-1]) */
    float* regressed_state;  /* [B,200,2] */
} gcpb200_seq_io;

const char* gcpb200_last_error(void);
const char* gcpb200_version(void);

int gcpb200_create(gcpb200_ctx** out, const gcpb200_config* cfg);
void gcpb200_destroy(gcpb200_ctx* ctx);

/* Packs the reference weights for the kernels: centre taps of the 3x3 "MLP" convs, eval-BatchNorm folded,
 * [W_ih|W_hh] concatenated and gate-interleaved, up-sample+pad+conv composites, bf16, K-major. */
int gcpb200_load_weights(gcpb200_ctx* ctx, const gcpb200_tensor* tensors, int n_tensors);

int gcpb200_rollout(gcpb200_ctx* ctx, const gcpb200_rollout_io* io, void* stream);

/* Sequential GCP rollout; the context must have been created with model = GCPB200_MODEL_SEQUENTIAL. */
int gcpb200_seq_rollout(gcpb200_ctx* ctx, const gcpb200_seq_io* io, void* stream);

/* L2 image cost over image sequences stored in time order, images [B,n_frames,3,32,32] (sequential model):
 * L2ImageCost._compute on rollouts[i][:end_ind[i]+1] (gcp/planning/cem/cost_fcn.py:9-22,65-72). */
int gcpb200_cost_l2_seq(gcpb200_ctx* ctx, const float* images, int n_frames, const int64_t* end_ind, const float* goal,
                        int B, int dense, float final_step_weight, float* cost /* [B] */, void* stream);

/* dst[c,t,:] = src[c, node_of_frame(c,t), :] for t <= end_ind[c], zeros after.  row_len % 4 == 0. */
int gcpb200_prune_gather(gcpb200_ctx* ctx, const float* src_df, const int64_t* end_ind, int B, int row_len,
                         float* dst /* [B,200,row_len] */, void* stream);

/* dst[c,t,:] = src_df[c, nodes[c,t], :] for t < len[c], zeros after: materialises outputs.pruned_prediction of the
 * adaptive model (adaptive.py:75) from the node lists gcpb200_rollout returned.  row_len % 4 == 0. */
int gcpb200_gather_nodes(gcpb200_ctx* ctx, const float* src_df, const int32_t* nodes /* [B,255] */,
                         const int32_t* len /* [B] */, int B, int row_len, float* dst /* [B,255,row_len] */, void* stream);

/* L2 image cost over the frames listed in nodes[c,:len[c]] (CostFcn.__call__ on pruned predictions). */
int gcpb200_cost_l2_nodes(gcpb200_ctx* ctx, const float* images_df, const int32_t* nodes, const int32_t* len,
                          const float* goal, int B, int dense, float final_step_weight, float* cost /* [B] */,
                          void* stream);

/* goal: [3,32,32] in [-1,1].  dense != 0: sum over frames, else last frame only. */
int gcpb200_cost_l2(gcpb200_ctx* ctx, const float* images_df, const int64_t* end_ind, const float* goal, int B,
                    int dense, float final_step_weight, float* cost /* [B] */, void* stream);

/* cost[c] = sum over consecutive pairs of cat(latents of c, goal_seq) of the learned pairwise cost. */
int gcpb200_cost_learned(gcpb200_ctx* ctx, const float* e_df, const int64_t* end_ind, int B,
                         const float* goal_seq /* [Lg,128] */, int Lg, float* cost /* [B] */, void* stream);

/* Learned pairwise cost on arbitrary row pairs of a latent table lat [R,128] (device):
 *   seg_off == NULL: cost[i] = cost_pred(cat(lat[idx1[i]], lat[idx2[i]])), i < n -- LearnedCostEstimate.__call__, ndarray
 *     branch (gcp/planning/cem/cost_fcn.py:84-87) with TestTimeCostModel.forward (cost_mdl.py:138-145), as the hierarchical
 *     optimiser calls it on (start, subgoal) and (subgoal, goal) latents (gcp/planning/tree_optimizer.py:96-99);
 *   seg_off != NULL (device int32[n_seg+1], ascending, seg_off[n_seg] <= n): cost[s] = sum of the pair costs
 *     seg_off[s] .. seg_off[s+1]-1 -- the list branch (cost_fcn.py:88-97) on segments (tree_optimizer.py:145-150).
 * idx1 / idx2: device int32[n].  n <= 200 * max_candidates. */
int gcpb200_cost_pairs(gcpb200_ctx* ctx, const float* lat, const int32_t* idx1, const int32_t* idx2, int n,
                       const int32_t* seg_off, int n_seg, float* cost, void* stream);

/* Closed-loop execution step (ImageCEMPolicy._infer_action, gcp/planning/planner_policy.py:215-221):
 * enc = encoder(img); action = inv_mdl.action_pred(cat(enc, target_latent)) (InverseModel.run_single,
 * gcp/prediction/models/auxilliary_models/inverse_mdl.py:221-224).  img [n,3,32,32] fp32 in [-1,1], target_latent
 * [n,128], action [n,2], enc [n,128] or NULL; all device pointers.  Uses the rollout workspace: stream-order it with
 * the rollouts of the same context. */
int gcpb200_infer_action(gcpb200_ctx* ctx, const float* img, const float* target_latent, int n, float* action /* [n,2] */,
                         float* enc /* [n,128] or NULL */, void* stream);

/* ---- training-phase forward + loss (BASELINE config 1: the 25-room prediction config, batch-statistic BatchNorm) ----
 * One call = `output = model(inputs); losses = model.loss(inputs, output); losses.total = model.get_total_loss(...)`
 * of the reference in .train() mode, forward only (no gradients: this is the validation pass of train.py:197-209).
 * The reference's three random draws are inputs, so the call is deterministic: the posterior noise `eps`
 * (q_z.sample(), blox/torch/dist.py:246-247), the inverse model's frame pair (inverse_mdl.py:88-98) and the cost model's
 * frame pair + target (cost_mdl.py:100-113; the target is the host-side cost function applied to the ground-truth frames).
 * Needs a GCPB200_MODEL_TREE context with attach_cost_mdl = 1 whose state dict held the training-only tensors
 * (inf_encoder.*, tree_module.tree_modules.k.inference.q.*), and B <= 128. */
#define GCPB200_LOSS_LEN_PRED 0
#define GCPB200_LOSS_ACTION_RECONST 1
#define GCPB200_LOSS_COST_ESTIMATION 2
#define GCPB200_LOSS_STATE_REGRESSION 3
#define GCPB200_LOSS_DENSE_IMG_REC 4
#define GCPB200_LOSS_KL 5
#define GCPB200_LOSS_EXISTENCE_PREDICTOR 6
#define GCPB200_LOSS_ENTROPY 7
#define GCPB200_LOSS_TOTAL 8
#define GCPB200_N_LOSSES 9
typedef struct {
    /* ---- inputs (device) ---- */
    const float* traj_seq;      /* [B,200,3,32,32] frames in [-1,1], zero past end_ind */
    const float* pad_mask;      /* [B,200] 1 = real frame */
    const int64_t* end_ind;     /* [B] index of the last real frame (>= 1) */
    const float* I_0;           /* [B,3,32,32] */
    const float* I_g;           /* [B,3,32,32] */
    const float* states;        /* [B,200,2] traj_seq_states */
    const float* actions;       /* [B,199,2] */
    const float* eps;           /* [B,255,256] N(0,1) posterior noise, depth-first node order */
    const int64_t* inv_t0;      /* [B] inverse-model frame pair: latent of frame t0 from the encoder, t1 from the tree */
    const int64_t* inv_t1;
    const int64_t* cost_start;  /* [B] cost-model pair (both from the tree's matched latents) */
    const int64_t* cost_end;
    const float* cost_target;   /* [B] ground-truth cost of that pair, or NULL: EuclideanPathLength of traj_seq[start..end]
                                   (cost_fcn.py:49-54, the 25-room config's cost_fcn) computed on the device */
    int B;
    /* ---- outputs (device; losses required, the rest may be NULL) ---- */
    float* losses;              /* [GCPB200_N_LOSSES] .value of each loss term, order above */
    float* nll_per_frame;       /* [B,200] reconstruction NLL summed over the frame's 3x32x32 sub-pixels, times pad_mask */
    float* kl_per_seq;          /* [B] */
    float* e_0;                 /* [B,128] */
    float* e_g;                 /* [B,128] */
    float* enc_traj_seq;        /* [B,200,128] */
    float* inf_enc_seq;         /* [B,200,128] */
    float* seq_len_logits;      /* [B,200] */
    float* e_df;                /* [B,255,128] node latents, depth-first */
    float* p_mu;                /* [B,255,256] prior */
    float* p_log_sigma;
    float* q_mu;                /* [B,255,256] approximate posterior */
    float* q_log_sigma;
    int32_t* match_timesteps;   /* [B,255] frame index every node is matched to */
    float* images_df;           /* [B,255,3,32,32] DLM mean image of every node (tree.df.images) */
    float* existence;           /* [B,255] */
    float* model_enc_seq;       /* [B,200,128] matched latents, zero padded */
    float* regressed_state;     /* [B,200,2] */
    float* inv_actions;         /* [B,2] */
    float* cost_pred;           /* [B] */
} gcpb200_train_io;
int gcpb200_forward_loss(gcpb200_ctx* ctx, const gcpb200_train_io* io, void* stream);

/* indices (and values) of the k lowest costs in ascending order; ties broken by index. */
int gcpb200_topk(gcpb200_ctx* ctx, const float* cost, int N, int k, int32_t* idx /* [k] */, float* val /* [k] or NULL */,
                 void* stream);

int gcpb200_refit(gcpb200_ctx* ctx, const float* z /* [N,255,256] */, const int32_t* elite_idx, int k,
                  float* mean /* [255*256] */, float* std /* [255*256] */, void* stream);

/* z[c] = clip(mean + std * n(seed, first_candidate_id + c)); mean/std NULL -> 0 / std_scalar. */
int gcpb200_sample_noise(gcpb200_ctx* ctx, const float* mean, const float* std, float std_scalar, uint64_t seed,
                         uint64_t first_candidate_id, int B, float clip, float* z /* [B,255,256] */, void* stream);

/* same, for an explicit list of global candidate ids (device int32[B]): regenerates e.g. the elites of a
 * sharded CEM iteration on every rank without moving samples between GPUs. */
int gcpb200_sample_noise_ids(gcpb200_ctx* ctx, const float* mean, const float* std, float std_scalar, uint64_t seed,
                             const int32_t* ids, int B, float clip, float* z /* [B,255,256] */, void* stream);

/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t gcpb200_launch_count(gcpb200_ctx* ctx);

/* Phase timing with CUDA events on the caller's stream (measurement aid for bench.py; off by default).
 * phases: 0 encoder+length, 1 tree recursion, 2 decoder GEMM layers, 3 decoder tail conv kernel, 4 heads,
 * 5 = the whole gcpb200_rollout call. */
#define GCPB200_N_PHASES 6
int gcpb200_profile_enable(gcpb200_ctx* ctx, int on);
int gcpb200_profile_read(gcpb200_ctx* ctx, double* ms /* [GCPB200_N_PHASES] */, int64_t* tail_images,
                         int64_t* tail_launches);

/* ---- DTW family.  No weights involved: any context of the device will do. ------------------------------
 *
 * gcpb200_cdist_mean   blox.torch.ops.batch_cdist(x, y, reduction='mean') (blox/torch/ops.py:62-91): the cost matrix of
 *                      AdaptiveBinding.get_w (gcp/prediction/models/adaptive_binding/adaptive.py:41-47) and of
 *                      DTWEvalBinding.get_single_matches (gcp/evaluation/evaluation_matching.py:138).
 *                      x [B,n,dim], y [B,m,dim] -> out [B,n,m] = mean_d (x - y)^2, fp32 (dim % 4 == 0, 16-byte aligned).
 * gcpb200_soft_dtw     soft_dtw(cost / temp, end_inds) (probabilistic_dtw.py:11-121; adaptive.py:51): float64 forward-backward
 *                      of the 'nohor' alignment.  cost [B,r,c] fp32 (r nodes >= c frames), end_inds [B] int64 or NULL (= c-1).
 *                      workspace: gcpb200_soft_dtw_workspace(B,r,c) bytes; on return it holds the forward table [B,r,c] then
 *                      the backward table [B,r,c] (float64, un-flipped coordinates).
 *                      w [B,r,c]: expected edge frequencies (the function's return value).
 *                      w_bf [B,r,c] or NULL: depthfirst2breadthfirst(normalize(w, 1)) = AdaptiveBinding.get_w's result
 *                      (adaptive.py:53-61; blox/torch/dist.py:22-24; gcp/prediction/utils/tree_utils.py:217-232), r = 2^d - 1.
 *                      rowsum_max [1] or NULL: max over (b, node) of sum_t w -- the reference warns and stops when this is
 *                      not within 1e-2 of 1 (probabilistic_dtw.py:115-117); the caller decides.
 * gcpb200_dtw          c_dtw / basic_dtw / batched_dtw (gcp/evaluation/dtw_utils.py:77-130; gcp/evaluation/cutils.pyx:21-28)
 *                      with _traceback / _batched_traceback (dtw_utils.py:201-241) and the per-frame best match of
 *                      get_single_matches (evaluation_matching.py:142-146).  cost [B,r,c] fp32 (cost_is_f64 = 0) or float64;
 *                      end_ind [B] int64 or NULL: column the path starts from (c-1).
 *                      acc [B,r+1,c+1] float64: the padded accumulated-cost table; dist [B] = acc[r][end+1] / (r+end+1);
 *                      path_p / path_q [B, r+c-1] int32: the warping path in start -> end order, RIGHT-aligned, the unused
 *                      leading slots hold (0,0) (as _batched_traceback pads); path_len [B];
 *                      match_inds [B,c] int32 or NULL: for every column the path row with the smallest accumulated cost
 *                      (first minimum; 0 for columns the path does not visit).
 * gcpb200_gather_rows  out[k] = src[idx[k]] for rows of row_floats floats (gen_images = estimates[inds],
 *                      evaluation_matching.py:147).
 */
int gcpb200_cdist_mean(gcpb200_ctx* ctx, const float* x, const float* y, int B, int n, int m, int dim, float* out, void* stream);
size_t gcpb200_soft_dtw_workspace(int B, int r, int c);
int gcpb200_soft_dtw(gcpb200_ctx* ctx, const float* cost, float temp, const int64_t* end_inds, int B, int r, int c,
                     void* workspace, float* w, float* w_bf, float* rowsum_max, void* stream);
int gcpb200_dtw(gcpb200_ctx* ctx, const void* cost, int cost_is_f64, const int64_t* end_ind, int B, int r, int c, double* acc,
                double* dist, int32_t* path_p, int32_t* path_q, int32_t* path_len, int32_t* match_inds, void* stream);
int gcpb200_gather_rows(gcpb200_ctx* ctx, const float* src, const int32_t* idx, int n_out, int row_floats, float* out,
                        void* stream);

/* ---- optimiser step (the `self.optimizer.step()` of train.py:162) -------------------------------------
 * The reference trains with RAdam (gcp_builder.py:174-186,259; blox/torch/radam.py:17-80) or torch.optim.Adam, wrapped in
 * ClipGradOptimizer (blox/torch/training.py:146-161: torch.nn.utils.clip_grad_norm_ over all parameters, then the step).
 *
 * gcpb200_sq_norm     *acc += sum(x[i]^2) in float64 (device scalar, zeroed by the caller): called once per gradient tensor,
 *                     it yields the total gradient norm the clipping needs without any host round trip.
 * gcpb200_optim_step  one in-place step over a flat fp32 parameter array p with gradient g and moments m, v (all [n]).
 *                     kind GCPB200_OPT_ADAM / GCPB200_OPT_RADAM; `step` is the 1-based step count of this update;
 *                     hyper-parameters are doubles (Python floats) and rounded to fp32 where torch rounds them.
 *                     grad_sq_norm (device float64 scalar or NULL) + max_norm > 0: gradients are scaled by
 *                     min(1, max_norm / (sqrt(*grad_sq_norm) + 1e-6)) as clip_grad_norm_ does.  HBM-bound: 16 B read and
 *                     12 B written per parameter. */
#define GCPB200_OPT_ADAM 0
#define GCPB200_OPT_RADAM 1
int gcpb200_sq_norm(gcpb200_ctx* ctx, const float* x, int64_t n, double* acc, void* stream);
int gcpb200_optim_step(gcpb200_ctx* ctx, int kind, float* p, const float* g, float* m, float* v, int64_t n, double lr,
                       double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       const double* grad_sq_norm, float max_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCPB200_H */

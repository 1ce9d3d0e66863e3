"""Host-side camera geometry feeding the epi module: pose files -> relative c2w / K ->
per-frame-pair fundamental matrices.  Mirrors ``animatediff/data/dataset_validation.py``
(calc_cross_product_matrix :28, calc_essential_matrix :45, calc_fundamental_matrix :57,
Camera :79, ValRealEstate10KPoseFolded.load_cameras_specific/get_batch :180-271) and the
per-UNet-call recompute of ``pipelines/pipeline_animation_epi_advanced.py:636-647``, which is
batched here instead of a Python loop over V*16 frames.  fp64 numpy, like the reference.
"""
import numpy as np
import torch


def calc_cross_product_matrix(vec):
    """[a]x for a 3-vector or an [N,3] array (dataset_validation.py:28-39)."""
    v = np.asarray(vec, dtype=np.float64)
    single = v.ndim == 1
    v = v.reshape(-1, 3)
    m = np.zeros((v.shape[0], 3, 3))
    m[:, 0, 1], m[:, 0, 2], m[:, 1, 2] = -v[:, 2], v[:, 1], -v[:, 0]
    m = m - m.transpose(0, 2, 1)
    return m[0] if single else m


def calc_essential_matrix(T_mat):
    """E = R [(-R^T t)]x; accepts [4,4] or a batch [...,4,4] (dataset_validation.py:45-50)."""
    T = np.asarray(T_mat, dtype=np.float64)
    R, t = T[..., :3, :3], T[..., :3, 3]
    t_ess = -np.einsum("...ji,...j->...i", R, t)
    cross = calc_cross_product_matrix(t_ess.reshape(-1, 3)).reshape(R.shape)
    return R @ cross


def calc_fundamental_matrix(T_mat, K_mat1, K_mat2):
    """F = K2^-T E K1^-1, batched over leading dims (dataset_validation.py:57-64)."""
    E = calc_essential_matrix(T_mat)
    K2_invT = np.swapaxes(np.linalg.inv(np.asarray(K_mat2, dtype=np.float64)), -1, -2)
    K1_inv = np.linalg.inv(np.asarray(K_mat1, dtype=np.float64))
    return K2_invT @ E @ K1_inv


class Camera:
    """One row of a RealEstate10K pose file (dataset_validation.py:79-91)."""

    def __init__(self, entry):
        self.cid = entry[0]
        self.fx, self.fy, self.cx, self.cy = entry[1:5]
        w2c = np.eye(4)
        w2c[:3, :] = np.array(entry[7:]).reshape(3, 4)
        self.w2c_mat = w2c
        self.c2w_mat = np.linalg.inv(w2c)


def read_pose_file(path):
    with open(path, "r") as fh:
        lines = fh.readlines()[1:]
    return [Camera([float(x) for x in ln.strip().split(" ")]) for ln in lines if ln.strip()]


def folded_pose_pair(pose_file_0, pose_file_1, sample_n_frames=16, sample_size=256):
    """F_mats float32 [2n,3,3], c2w [2n-1,4,4], K [2n-1,3,3] of the folded two-trajectory layout
    that inference_epi.py feeds the pipeline (dataset_validation.py:180-271, quirks included:
    H,W = 1280,720; trajectory 1 reversed; K of trajectory 0 used for both)."""
    cams0 = read_pose_file(pose_file_0)
    cams1 = read_pose_file(pose_file_1)[::-1]
    H, W = 1280, 720
    crop = min(H, W)
    rescale = sample_size / crop
    dH, dW = (H - crop) / 2, (W - crop) / 2
    K0 = np.array([[[W * rescale * c.fx, 0, (W * c.cx - dW) * rescale],
                    [0, H * rescale * c.fy, (H * c.cy - dH) * rescale],
                    [0, 0, 1]] for c in cams0])

    def relative(cams):
        inv0 = np.linalg.inv(cams[0].c2w_mat)
        return np.array([inv0 @ c.c2w_mat for c in cams], dtype=np.float32)

    c2w = np.concatenate([relative(cams0)[1:][::-1], relative(cams1)], axis=0)
    K = np.concatenate([K0[1:][::-1], K0], axis=0)
    n = sample_n_frames
    sid = n - 1 - np.arange(n)
    tid = n - 1 + np.arange(n)
    s2t = np.linalg.inv(c2w[tid].astype(np.float64)) @ c2w[sid]
    Fm = torch.as_tensor(calc_fundamental_matrix(s2t, K[sid], K[tid])).float()
    return torch.cat([Fm, Fm.permute(0, 2, 1)], dim=0).contiguous(), c2w, K


def folded_pose_conditioning(pose_file_0, pose_file_1, sample_n_frames=16, sample_size=256):
    """Everything ``ValRealEstate10KPoseFolded.get_batch`` derives from the two pose files (dataset_validation.py:180-271):
    ``F_mats`` float32 [2n,3,3]; the FOLDED camera list (frame i of video 0 = camera n-1-i, of video 1 = camera n-1+i,
    :262-268) as ``c2w`` float32 [2n,4,4], ``K`` [2n,3,3] and ``intrinsics`` float32 [2n,4] = (fx, fy, cx, cy) in pixels of
    the ``sample_size`` image -- the input of the Pluecker embedding (``ops.plucker_embedding``), which the pipeline
    then reshapes to ``[2, 6, n, H, W]`` (inference_epi.py:252-254).  Quirks kept: H, W = 1280, 720 swapped; trajectory 1
    reversed and made relative to ITS first (reversed) camera; K of trajectory 0 for both, but the intrinsics of
    trajectory 1 for the second half (:236-238)."""
    cams0 = read_pose_file(pose_file_0)
    cams1 = read_pose_file(pose_file_1)[::-1]
    H, W = 1280, 720
    crop = min(H, W)
    rescale = sample_size / crop
    dH, dW = (H - crop) / 2, (W - crop) / 2

    def intr(c):
        return [W * rescale * c.fx, H * rescale * c.fy, (W * c.cx - dW) * rescale, (H * c.cy - dH) * rescale]
    i0 = np.array([intr(c) for c in cams0])
    i1 = np.array([intr(c) for c in cams1])
    Fm, c2w, K = folded_pose_pair(pose_file_0, pose_file_1, sample_n_frames, sample_size)
    intrinsics = np.concatenate([i0[1:][::-1], i1], axis=0)
    n = sample_n_frames
    fold = np.concatenate([n - 1 - np.arange(n), n - 1 + np.arange(n)])
    return dict(F_mats=Fm, c2w=torch.from_numpy(np.ascontiguousarray(c2w[fold])).float(),
                K=np.ascontiguousarray(K[fold]), intrinsics=torch.from_numpy(np.ascontiguousarray(intrinsics[fold])).float(),
                fold_indices=fold, c2w_all=c2w, intrinsics_all=intrinsics)


def sample_clip_frames(total_frames, sample_n_frames=16, sample_stride=8, rng=None):
    """The 2n-1 frame numbers the TRAINING dataset draws from a clip (dataset_train_realestate10k.py:399-406): an evenly
    strided window at a random start (ONE ``random.randint`` from Python's global RNG, like the reference)."""
    import random
    rng = random if rng is None else rng
    sample_length = sample_n_frames * 2 - 1
    assert total_frames >= sample_length
    stride = min(total_frames // sample_length, sample_stride)
    clip_length = min(total_frames, (sample_length - 1) * stride + 1)
    start = rng.randint(0, total_frames - clip_length)
    return np.linspace(start, start + clip_length - 1, sample_length, dtype=int)


def folded_clip_conditioning(pose_file, frame_indices, image_hw, sample_n_frames=16, sample_size=256):
    """Everything ``RealEstate10KPoseFolded.get_batch`` of the TRAINING dataset derives from the camera file of one clip
    (dataset_train_realestate10k.py:408-474): the 2n-1 cameras ``frame_indices`` made relative to the middle one (float32,
    :437), pixel intrinsics of the centre-cropped ``sample_size`` image from the clip's real image size ``image_hw``
    (:421-425), F of frame n-1-i -> frame n-1+i and its transpose for the way back (:451-462), and the FOLDED camera list
    (video 0 runs from the middle frame backwards, video 1 forwards, :463-469) whose Pluecker rays condition the two videos.
    Returns the same dictionary as :func:`folded_pose_conditioning` (``F_mats`` [2n,3,3], ``c2w`` [2n,4,4], ``K`` [2n,3,3],
    ``intrinsics`` [2n,4])."""
    cams = read_pose_file(pose_file)
    cams = [cams[int(i)] for i in frame_indices]
    n = sample_n_frames
    assert len(cams) == 2 * n - 1
    H, W = image_hw
    crop = min(H, W)
    rescale = sample_size / crop
    dH, dW = (H - crop) / 2, (W - crop) / 2
    K = np.array([[[W * rescale * c.fx, 0, (W * c.cx - dW) * rescale],
                   [0, H * rescale * c.fy, (H * c.cy - dH) * rescale],
                   [0, 0, 1]] for c in cams])
    intrinsics = np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], axis=1)
    inv_mid = np.linalg.inv(cams[n - 1].c2w_mat)
    c2w = np.array([inv_mid @ c.c2w_mat for c in cams], dtype=np.float32)
    sid = n - 1 - np.arange(n)
    tid = n - 1 + np.arange(n)
    s2t = np.linalg.inv(c2w[tid].astype(np.float64)) @ c2w[sid]
    Fm = torch.as_tensor(calc_fundamental_matrix(s2t, K[sid], K[tid])).float()
    fold = np.concatenate([sid, tid])
    return dict(F_mats=torch.cat([Fm, Fm.permute(0, 2, 1)], dim=0).contiguous(),
                c2w=torch.from_numpy(np.ascontiguousarray(c2w[fold])).float(), K=np.ascontiguousarray(K[fold]),
                intrinsics=torch.from_numpy(np.ascontiguousarray(intrinsics[fold])).float(), fold_indices=fold,
                c2w_all=c2w, intrinsics_all=intrinsics)


def interpolate_view_poses(src_poses, tgt_poses, split_num):
    """``RealEstate10KPoseFolded.interpolate_poses`` (dataset_train_realestate10k.py:365-384): ``split_num`` trajectories
    between two n-frame trajectories [n,4,4] -- translation blended linearly, rotation by spherical interpolation, per
    frame -- stacked view-major [split_num * n, 4, 4] (view 0 = ``src_poses``, view split_num - 1 = ``tgt_poses``)."""
    from scipy.spatial.transform import Rotation, Slerp
    src = np.asarray(src_poses)
    tgt = np.asarray(tgt_poses)
    n = len(src)
    out = np.tile(src, (split_num, 1, 1))
    w = np.linspace(0.0, 1.0, split_num)
    for i in range(split_num):
        alpha = i / (split_num - 1)
        out[i * n:(i + 1) * n, :3, 3] = src[:, :3, 3] * (1 - alpha) + tgt[:, :3, 3] * alpha
    for f in range(n):
        ends = Rotation.from_matrix(np.stack([src[f, :3, :3], tgt[f, :3, :3]]))
        out[f::n, :3, :3] = Slerp([0, 1], ends)(w).as_matrix()
    return out


def multiview_clip_conditioning(cond, split_num):
    """The ``validation_video_split != 2`` branch of the training dataset (dataset_train_realestate10k.py:466-478): from the
    folded two-video conditioning of :func:`folded_clip_conditioning`, ``split_num`` views interpolated between video 0
    and video 1 -- cameras ``c2w`` float32 [split_num n, 4, 4] (what the V-view pipeline receives as ``aux_c2w``), ``K``
    [split_num n, 3, 3] and ``intrinsics`` [split_num n, 4] (video 0's, repeated per view), the inputs of the Pluecker rays."""
    n = cond["c2w"].shape[0] // 2
    c2w = cond["c2w"].numpy()
    inter = interpolate_view_poses(c2w[:n], c2w[n:], split_num)
    intr = cond["intrinsics"][:n].repeat(split_num, 1)
    return dict(c2w=torch.from_numpy(np.ascontiguousarray(inter)).float(), K=np.tile(cond["K"][:n], (split_num, 1, 1)),
                intrinsics=intr.contiguous())


def pairwise_fundamental_matrices(c2w, K, video_id, n_frames=16):
    """Per-UNet-call F for a view pairing (pipeline_animation_epi_advanced.py:636-647):
    row (v, f) gets F of source (v, f) -> target (video_id[v], f).  c2w [V*n,4,4], K [V*n,3,3]."""
    c2w = np.asarray(c2w, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    V = len(video_id)
    src = np.arange(V * n_frames)
    tgt = (np.asarray(video_id)[:, None] * n_frames + np.arange(n_frames)[None]).reshape(-1)
    s2t = np.linalg.inv(c2w[tgt]) @ c2w[src]
    return torch.as_tensor(calc_fundamental_matrix(s2t, K[src], K[tgt])).float()

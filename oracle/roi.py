"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the reference's RoI preparation, the checker for givepose_b200/roi.py.

Follows ``tools/dataset_utils.py`` (``get_2d_coord_np`` :8-30, ``crop_resize_by_warp_affine`` :101-114, ``get_affine_transform``
:116-157, ``get_dir`` :159-166, ``get_3rd_point`` :168-170) and the call sites ``evaluation/load_data_eval.py:256-289``.
The pixel arithmetic itself lives in a third-party dependency that is NOT vendored under /root/reference:
opencv-python==4.8.0.76 (GIVEPose_env.yml:250).  Its published algorithm is restated here:
  * ``cv::getAffineTransform`` (imgwarp.cpp): the 6x6 system ``[x y 1 0 0 0; 0 0 0 x y 1] X = [u; v]`` solved by ``cv::solve``
    (DECOMP_LU -> ``LUImpl<double>``, matrix_decomp.cpp: partial pivoting, ``alpha = A[j][i] * (-1/A[i][i])``);
  * ``cv::warpAffine`` (imgwarp.cpp): inversion of the 2x3 matrix, then for INTER_NEAREST the 10-bit fixed-point source index
    ``X = (cvRound((M1*y + M2)*1024) + 512 + cvRound(M0*x*1024)) >> 10`` clamped to int16, BORDER_CONSTANT 0.
Pinned: ``tests/golden/roi.npz`` holds outputs of the reference's own functions run on the cv2 of the build image
(``tests/golden/make_golden_roi.py``); the restatement is bit-identical to them and to live ``cv2.warpAffine`` / ``cv2.getAffineTransform``
calls (tests/test_roi_oracle.py).
"""
import numpy as np


def get_2d_coord_np(width, height):
    x = np.linspace(0, width - 1, width, dtype=np.float32)
    y = np.linspace(0, height - 1, height, dtype=np.float32)
    x = (x - np.float32((width - 1) / 2)) / np.float32((width - 1) / 2)
    y = (y - np.float32((height - 1) / 2)) / np.float32((height - 1) / 2)
    return np.asarray(np.meshgrid(x, y))   # (2, H, W)


def affine_points(center, scale, out):
    """The three point pairs of get_affine_transform(center, (scale, scale), 0, (out, out)); float32 storage."""
    center = np.asarray(center, dtype=np.float64)
    scale_tmp = (float(scale), float(scale))
    shift = np.array([0, 0], dtype=np.float32)
    src_w, dst_w, dst_h = scale_tmp[0], out, out
    sn, cs = np.sin(0.0), np.cos(0.0)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    third = lambda a, b: b + np.array([-(a - b)[1], (a - b)[0]], dtype=np.float32)
    src[2:, :] = third(src[0, :], src[1, :])
    dst[2:, :] = third(dst[0, :], dst[1, :])
    return src, dst


def get_affine_transform_cv(src, dst):
    """cv::getAffineTransform restated (LUImpl<double>)."""
    m = 6
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[2 * i, 0] = A[2 * i + 1, 3] = float(src[i, 0])
        A[2 * i, 1] = A[2 * i + 1, 4] = float(src[i, 1])
        A[2 * i, 2] = A[2 * i + 1, 5] = 1.0
        b[2 * i], b[2 * i + 1] = float(dst[i, 0]), float(dst[i, 1])
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < np.finfo(np.float64).eps * 100:
            raise ValueError("singular RoI transform")
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, m):
            alpha = A[j, i] * d
            for kk in range(i + 1, m):
                A[j, kk] += alpha * A[i, kk]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for k in range(i + 1, m):
            s -= A[i, k] * b[k]
        b[i] = s / A[i, i]
    return b.reshape(2, 3)


def invert_affine(Mf):
    M = np.array(Mf, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def affine_inverse(center, scale, out):
    return invert_affine(get_affine_transform_cv(*affine_points(center, scale, out)))


def source_index(Minv, out_w, out_h):
    """(Y, X) int arrays (out_h, out_w) of cv::warpAffine INTER_NEAREST."""
    AB = 1024.0
    x = np.arange(out_w, dtype=np.float64)
    y = np.arange(out_h, dtype=np.float64)
    ad = np.rint(Minv[0] * x * AB).astype(np.int64)
    bd = np.rint(Minv[3] * x * AB).astype(np.int64)
    X0 = np.rint((Minv[1] * y + Minv[2]) * AB).astype(np.int64) + 512
    Y0 = np.rint((Minv[4] * y + Minv[5]) * AB).astype(np.int64) + 512
    X = np.clip((X0[:, None] + ad[None, :]) >> 10, -32768, 32767)
    Y = np.clip((Y0[:, None] + bd[None, :]) >> 10, -32768, 32767)
    return Y, X


def warp_nearest(img, Minv, out):
    Y, X = source_index(Minv, out, out)
    H, W = img.shape[:2]
    ok = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
    res = np.zeros((out, out) + img.shape[2:], dtype=img.dtype)
    res[ok] = img[Y[ok], X[ok]]
    return res


def crop_resize_nearest(img, center, scale, out):
    """crop_resize_by_warp_affine(img, center, scale, out, interpolation=cv2.INTER_NEAREST)."""
    return warp_nearest(img, affine_inverse(center, scale, out), out)


def roi_tensors(image, mask, bbox_center, img_scale, inst_id=-1, img_size=256, out_res=64, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """One RoI as load_data_eval.py:256-289 prepares it: roi_img (3,S,S), roi_mask (1,S,S), roi_coord_2d (2,R,R), float32."""
    H, W = image.shape[:2]
    roi_img = crop_resize_nearest(image, bbox_center, img_scale, img_size)
    roi_img = ((roi_img / 255.0 - mean) / std).transpose(2, 0, 1).astype(np.float32)
    coord = get_2d_coord_np(W, H).transpose(1, 2, 0)
    roi_coord = crop_resize_nearest(coord, bbox_center, img_scale, out_res).transpose(2, 0, 1)
    mt = mask.astype(np.float32) if inst_id < 0 else (mask == inst_id).astype(np.float32)
    roi_mask = crop_resize_nearest(mt, bbox_center, img_scale, img_size)[None]
    return roi_img, roi_mask, roi_coord.astype(np.float32)


def resize_linear_u8(img, dw, dh):
    """cv2.resize(img, (dw, dh)) for uint8 images, INTER_LINEAR: OpenCV's fixed-point path (resize.cpp: INTER_RESIZE_COEF_BITS = 11,
    HResizeLinear + VResizeLinear<uchar>), exact for downscaling / same size (the loaders' 480x640 -> 256x256, load_data_eval.py:336)."""
    def coeffs(n_dst, n_src):
        scale = 1.0 / (np.float64(n_dst) / np.float64(n_src))
        idx = np.zeros(n_dst, np.int64)
        a = np.zeros((n_dst, 2), np.int64)
        for d in range(n_dst):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if s < 0:
                f, s = np.float32(0), 0
            if s >= n_src - 1:
                f, s = np.float32(0), n_src - 1
            idx[d] = s
            a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        return idx, a
    H, W = img.shape[:2]
    xi, xa = coeffs(dw, W)
    yi, ya = coeffs(dh, H)
    src = img.astype(np.int64)
    x1, y1 = np.minimum(xi + 1, W - 1), np.minimum(yi + 1, H - 1)
    rows = src[:, xi, :] * xa[None, :, 0, None] + src[:, x1, :] * xa[None, :, 1, None]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (rows[yi] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def full_img(image, resize=(256, 256), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """load_data_eval.py:336-338: optional cv2.resize, normalisation, HWC -> CHW (float32 at tensor creation, :364)."""
    im = resize_linear_u8(image, *resize) if resize is not None else image
    return ((im / 255.0 - mean) / std).transpose(2, 0, 1).astype(np.float32)

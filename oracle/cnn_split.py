"""Stand-in oracle of the CNN CU-split predictor.  TEST INFRASTRUCTURE - PARITY UNPINNED.

The reference holds no inference code, no weights and no encoder hook for this predictor: only the
training-set dumper (/root/reference/Src_HARP/CShow_PredResiReco.h:127-255) and the MATLAB layer list
(/root/reference/matlab/train/TrainABasicConvolutionalNeuralNetworkForClassificationExample.m:70-91).  This file
therefore *constructs* the fp32 predictor BASELINE.json configs[4] calls "the reference OpenCV fp32 predictor":

  input   32x32x3 uint8, the ORIGINAL luma of a depth-1 CU replicated over three channels
          (CShow_PredResiReco.h:150,158,252; CHelper.h:267-280), minus the zero-centre image of MATLAB's
          imageInputLayer (identical for the three channels, because the channels are)
  conv 3x3x16 pad 1 - BN - ReLU - maxpool 2/2
  conv 3x3x32 pad 1 - BN - ReLU - maxpool 2/2
  conv 3x3x64 pad 1 - BN - ReLU
  FC 2 (features flattened channel-major: c, y, x) - softmax            (...Example.m:70-91)
  label 1 = "the depth-1 CU is split further" (matlab/.../detectAndClassify32Cu.m:57-62)

forward() is plain numpy fp32.  darknet_files() writes the same network as a Darknet cfg + weights pair, which
cv2.dnn.readNetFromDarknet runs in fp32: that OpenCV run is what pins forward() (tests/test_cnn_split.py; this image has
cv2.dnn but no onnx package, so the ONNX route SURVEY 8c proposes is not available).  Weights are seeded (make_weights):
nothing trained exists.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import struct

import numpy as np

CHANNELS = [(3, 16), (16, 32), (32, 64)]
BN_EPS = 1e-5          # MATLAB batchNormalizationLayer default
DARKNET_BN_EPS = 1e-6  # what OpenCV's Darknet importer bakes in


def make_weights(seed=20240):
    """Seeded fp32 weights in PyTorch / MATLAB-like layouts: conv [Cout][Cin][3][3], fc [2][64*8*8] (c, y, x)."""
    rng = np.random.default_rng(seed)
    w = {"bn_eps": np.float32(BN_EPS)}
    # the zero-centre image: smooth, around mid grey, like the mean of a training set of luma blocks
    yy, xx = np.mgrid[0:32, 0:32]
    w["mean"] = (118.0 + 6.0 * np.sin(xx / 9.0) + 4.0 * np.cos(yy / 7.0)).astype(np.float32)
    for i, (cin, cout) in enumerate(CHANNELS):
        fan_in = cin * 9
        w["conv%d_w" % i] = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        w["conv%d_b" % i] = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        w["bn%d_gamma" % i] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
        w["bn%d_beta" % i] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        # running statistics of the layer's pre-activation: scaled to what the layer really sees, so that the
        # normalised activations stay O(1) as they do in a trained network
        w["bn%d_mean" % i] = np.zeros(cout, np.float32)
        w["bn%d_var" % i] = np.ones(cout, np.float32)
    w["fc_w"] = (rng.standard_normal((2, 64 * 8 * 8)) * np.sqrt(1.0 / 4096)).astype(np.float32)
    w["fc_b"] = np.array([0.0, 0.0], np.float32)
    _calibrate(w, rng)
    return w


def _calibrate(w, rng):
    """Sets the BN running statistics from a calibration batch (what training would have left behind) and centres
    the two logits, so that the seeded network splits about half of natural-looking CUs."""
    n = 256
    yy, xx = np.mgrid[0:32, 0:32]
    patches = np.empty((n, 32, 32), np.uint8)
    for k in range(n):
        a, b, c = rng.uniform(0, 50), rng.uniform(3, 30), rng.uniform(0, 14)
        img = 128 + a * np.sin((xx + 3 * k) / b) + 0.6 * a * np.cos((yy - 2 * k) / (b * 0.7)) + rng.normal(0, c, (32, 32))
        if k % 3 == 0:
            img[rng.integers(4, 28):, rng.integers(4, 28):] -= rng.uniform(20, 70)
        patches[k] = np.clip(img, 0, 255)
    x = _input_blob(w, patches)
    for i in range(3):
        z = _conv3x3(x, w["conv%d_w" % i], w["conv%d_b" % i])
        w["bn%d_mean" % i] = z.mean(axis=(0, 2, 3)).astype(np.float32)
        w["bn%d_var" % i] = z.var(axis=(0, 2, 3)).astype(np.float32)
        x = np.maximum(_bn(w, i, z), 0)
        if i < 2:
            x = _maxpool2(x)
    lg = x.reshape(n, -1) @ w["fc_w"].T
    w["fc_b"] = np.array([0.0, -float(np.median(lg[:, 1] - lg[:, 0]))], np.float32)


def _input_blob(w, patches_u8):
    p = np.asarray(patches_u8)
    assert p.ndim == 3 and p.shape[1:] == (32, 32)
    x = p.astype(np.float32) - w["mean"][None]
    return np.repeat(x[:, None], 3, axis=1)                       # luma replicated x3 (CShow_PredResiReco.h:150)


def _conv3x3(x, wt, b):
    n, cin, h, ww = x.shape
    xp = np.zeros((n, cin, h + 2, ww + 2), np.float32)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((n, cin, 3, 3, h, ww), np.float32)
    for dy in range(3):
        for dx in range(3):
            cols[:, :, dy, dx] = xp[:, :, dy:dy + h, dx:dx + ww]
    out = np.einsum("ok,nkp->nop", wt.reshape(wt.shape[0], -1), cols.reshape(n, cin * 9, h * ww), dtype=np.float32)
    return (out + b[None, :, None]).reshape(n, wt.shape[0], h, ww).astype(np.float32)


def _bn(w, i, z):
    s = w["bn%d_gamma" % i] / np.sqrt(w["bn%d_var" % i] + w["bn_eps"])
    return ((z - w["bn%d_mean" % i][None, :, None, None]) * s[None, :, None, None]
            + w["bn%d_beta" % i][None, :, None, None]).astype(np.float32)


def _maxpool2(x):
    n, c, h, ww = x.shape
    return x.reshape(n, c, h // 2, 2, ww // 2, 2).max(axis=(3, 5))


def forward(w, patches_u8):
    """[N][32][32] uint8 luma -> fp32 logits [N][2] (before the softmax; split <=> logits[:,1] > logits[:,0])."""
    x = _input_blob(w, patches_u8)
    for i in range(3):
        x = np.maximum(_bn(w, i, _conv3x3(x, w["conv%d_w" % i], w["conv%d_b" % i])), 0)
        if i < 2:
            x = _maxpool2(x)
    return (x.reshape(x.shape[0], -1) @ w["fc_w"].T + w["fc_b"][None]).astype(np.float32)


def split_decision(logits):
    return (logits[:, 1] > logits[:, 0]).astype(np.uint8)


def cu_positions(width, height):
    """Depth-1 CUs (32x32) that lie completely inside the picture: CTUs in raster order, the four CUs of a CTU in
    z-order.  CUs cut by the picture edge are split by HM without any decision (TEncCu.cpp:519-521 bBoundary)."""
    pos = []
    for cy in range(0, height, 64):
        for cx in range(0, width, 64):
            for k in range(4):
                x, y = cx + 32 * (k & 1), cy + 32 * (k >> 1)
                if x + 32 <= width and y + 32 <= height:
                    pos.append((x, y))
    return np.array(pos, np.int32).reshape(-1, 2)


def gather_patches(luma, bit_depth=8):
    """Luma plane [H][W] -> uint8 patches of cu_positions() (samples above 8 bits are reduced by >> (bitDepth-8),
    as the reference's 8-bit PNG dump does for its training set)."""
    h, ww = luma.shape
    pos = cu_positions(ww, h)
    l8 = (np.asarray(luma).astype(np.int32) >> (bit_depth - 8)).astype(np.uint8)
    return np.stack([l8[y:y + 32, x:x + 32] for x, y in pos]) if len(pos) else np.zeros((0, 32, 32), np.uint8)


def darknet_files(w, cfg_path, weights_path):
    """The same network as a Darknet cfg/weights pair for cv2.dnn.readNetFromDarknet (OpenCV fp32 run = the pin)."""
    cfg = ["[net]", "width=32", "height=32", "channels=3", ""]
    for i, (cin, cout) in enumerate(CHANNELS):
        cfg += ["[convolutional]", "batch_normalize=1", "filters=%d" % cout, "size=3", "stride=1", "pad=1",
                "activation=relu", ""]
        if i < 2:
            cfg += ["[maxpool]", "size=2", "stride=2", ""]
    cfg += ["[connected]", "output=2", "activation=linear", ""]
    with open(cfg_path, "w") as f:
        f.write("\n".join(cfg))
    with open(weights_path, "wb") as f:
        f.write(struct.pack("<iii", 0, 2, 0) + struct.pack("<Q", 0))     # major, minor, revision, seen
        for i in range(3):
            # Darknet's batch-normalised convolution has no bias of its own: the conv bias moves into the mean;
            # the variance is shifted so that OpenCV's fixed epsilon gives MATLAB's
            f.write(w["bn%d_beta" % i].astype("<f4").tobytes())
            f.write(w["bn%d_gamma" % i].astype("<f4").tobytes())
            f.write((w["bn%d_mean" % i] - w["conv%d_b" % i]).astype("<f4").tobytes())
            f.write((w["bn%d_var" % i] + (w["bn_eps"] - np.float32(DARKNET_BN_EPS))).astype("<f4").tobytes())
            f.write(w["conv%d_w" % i].astype("<f4").tobytes())
        f.write(w["fc_b"].astype("<f4").tobytes())
        f.write(w["fc_w"].astype("<f4").tobytes())


def forward_opencv(w, patches_u8, tmpdir):
    """Runs the Darknet export through OpenCV's fp32 dnn module."""
    import os

    import cv2
    cfg, wts = os.path.join(tmpdir, "split32.cfg"), os.path.join(tmpdir, "split32.weights")
    darknet_files(w, cfg, wts)
    net = cv2.dnn.readNetFromDarknet(cfg, wts)
    net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
    net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
    net.setInput(_input_blob(w, patches_u8))
    return net.forward().reshape(len(patches_u8), 2).astype(np.float32)


def forward_fp16_model(w, patches_u8, stages=None):
    """Numerical model of the CUDA kernel (fasthevc_b200/csrc/fhc_cnn.cu): fp16 operands (the BN scale folded into
    the fp16 weights, the three replicated input channels summed into one), fp32 accumulation, fp16 activations
    between the layers, fp32 features into an fp32 FC.  Used to choose the tolerance the GPU test states.
    `stages` (a list) receives the pooled conv1 map, the pooled conv2 map and the conv3 features ([N][C][H][W])."""
    f = fold_for_kernel(w)
    p = np.asarray(patches_u8).astype(np.float32) - w["mean"][None]
    x = p.astype(np.float16).astype(np.float32)[:, None]
    for i in range(3):
        z = _conv3x3(x, f["w%d" % i].astype(np.float32), np.zeros(f["w%d" % i].shape[0], np.float32))
        x = np.maximum(z + f["t%d" % i][None, :, None, None], 0)
        if i < 2:
            x = _maxpool2(x.astype(np.float16).astype(np.float32))
        if stages is not None:
            stages.append(x)
    return (x.reshape(x.shape[0], -1) @ w["fc_w"].T + w["fc_b"][None]).astype(np.float32)


def fold_for_kernel(w):
    """BN folded: y = conv(x, w * s) + t with s = gamma / sqrt(var + eps), t = beta + s * (b - mean); layer 0 also
    sums the three identical input channels.  Weights rounded to fp16 (what the tensor cores read), t stays fp32."""
    f = {}
    for i in range(3):
        s = w["bn%d_gamma" % i] / np.sqrt(w["bn%d_var" % i] + w["bn_eps"])
        wt = w["conv%d_w" % i] * s[:, None, None, None]
        if i == 0:
            wt = wt.sum(axis=1, keepdims=True)
        f["w%d" % i] = wt.astype(np.float16)
        f["t%d" % i] = (w["bn%d_beta" % i] + s * (w["conv%d_b" % i] - w["bn%d_mean" % i])).astype(np.float32)
    return f

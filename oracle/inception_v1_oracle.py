"""ORACLE (test infrastructure, not product code) -- InceptionV1 encoder forward.

NumPy restatement of the reference's encoder graph:
  common/nets/inception_v1.py:29-339 (layers), common/nets/inception_utils.py:
  32-82 (arg scope: conv without bias -> slim.batch_norm(scale=False,
  epsilon=1e-3) -> ReLU), called with is_training=False from
  src/model_base.py:71-77, head + feature-map reshape src/model_base.py:79-104.

The arithmetic lives in TensorFlow 1.9 `tf.contrib.slim` (third party, absent
from /root/reference); its published semantics are restated here:
  * conv2d/max_pool2d `SAME`: out = ceil(in/s); pad = max((out-1)s+k-in, 0);
    pad_before = pad//2 (extra pixel at bottom/right); max-pool ignores pads.
  * batch_norm inference: (x - moving_mean) * rsqrt(moving_var + eps) + beta.
  * avg_pool2d 7x7 stride 1 VALID.

PARITY UNPINNED for numerics (the reference's tests hold no golden values);
pinned by the reference's known answers only: end-point shapes
(common/nets/inception_v1_test.py:98-115) and the parameter count 5,607,184
(common/nets/inception_v1_test.py:124-132) -- see tests/test_oracle_known_answers.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import numpy as np

CNN = 'Model/encoder/cnn/InceptionV1/'


def _same_pads(n, k, s):
    out = -(-n // s)
    pad = max((out - 1) * s + k - n, 0)
    return out, pad // 2, pad - pad // 2


def conv2d_same(x, w, stride):
    """x [B,H,W,Cin] NHWC, w [kh,kw,Cin,Cout] HWIO -> [B,Ho,Wo,Cout]."""
    B, H, W_, Cin = x.shape
    kh, kw, _, Cout = w.shape
    Ho, pt, pb = _same_pads(H, kh, stride)
    Wo, pl, pr = _same_pads(W_, kw, stride)
    if kh == 1 and kw == 1 and stride == 1:
        return (x.reshape(-1, Cin) @ w.reshape(Cin, Cout)).reshape(B, H, W_, Cout)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B * Ho * Wo, Cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (Ho - 1) * stride + 1:stride,
                       j:j + (Wo - 1) * stride + 1:stride, :]
            out += patch.reshape(-1, Cin) @ w[i, j]
    return out.reshape(B, Ho, Wo, Cout)


def max_pool_same(x, k, stride):
    B, H, W_, C = x.shape
    Ho, pt, pb = _same_pads(H, k, stride)
    Wo, pl, pr = _same_pads(W_, k, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), constant_values=-np.inf)
    out = np.full((B, Ho, Wo, C), -np.inf, dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            out = np.maximum(out, xp[:, i:i + (Ho - 1) * stride + 1:stride,
                                     j:j + (Wo - 1) * stride + 1:stride, :])
    return out


def conv_bn_relu(x, W, scope, stride=1, eps=1e-3):
    """slim.conv2d under inception_arg_scope (inception_utils.py:56-82)."""
    p = CNN + scope
    y = conv2d_same(x, W[p + '/weights'], stride)
    dt = x.dtype
    mean = W[p + '/BatchNorm/moving_mean'].astype(dt)
    var = W[p + '/BatchNorm/moving_variance'].astype(dt)
    beta = W[p + '/BatchNorm/beta'].astype(dt)
    inv = (1.0 / np.sqrt(var + dt.type(eps))).astype(dt)
    y = (y - mean) * inv + beta
    return np.maximum(y, 0)


def _mixed(x, W, name):
    """One inception block, branches concatenated [B0,B1,B2,B3] on channels
    (e.g. inception_v1.py:95-111)."""
    b2b = 'Conv2d_0a_3x3' if name == 'Mixed_5b' else 'Conv2d_0b_3x3'   # :240
    b0 = conv_bn_relu(x, W, name + '/Branch_0/Conv2d_0a_1x1')
    b1 = conv_bn_relu(x, W, name + '/Branch_1/Conv2d_0a_1x1')
    b1 = conv_bn_relu(b1, W, name + '/Branch_1/Conv2d_0b_3x3')
    b2 = conv_bn_relu(x, W, name + '/Branch_2/Conv2d_0a_1x1')
    b2 = conv_bn_relu(b2, W, name + '/Branch_2/' + b2b)
    b3 = max_pool_same(x, 3, 1)
    b3 = conv_bn_relu(b3, W, name + '/Branch_3/Conv2d_0b_1x1')
    return np.concatenate([b0, b1, b2, b3], axis=3)


def inception_v1(images, W):
    """inception_v1(num_classes=None, is_training=False): returns
    (net [B,1,1,1024], end_points).  inception_v1.py:269-339 (early return
    :328-329)."""
    ep = {}
    net = conv_bn_relu(images, W, 'Conv2d_1a_7x7', stride=2); ep['Conv2d_1a_7x7'] = net
    net = max_pool_same(net, 3, 2); ep['MaxPool_2a_3x3'] = net
    net = conv_bn_relu(net, W, 'Conv2d_2b_1x1'); ep['Conv2d_2b_1x1'] = net
    net = conv_bn_relu(net, W, 'Conv2d_2c_3x3'); ep['Conv2d_2c_3x3'] = net
    net = max_pool_same(net, 3, 2); ep['MaxPool_3a_3x3'] = net
    for name in ['Mixed_3b', 'Mixed_3c']:
        net = _mixed(net, W, name); ep[name] = net
    net = max_pool_same(net, 3, 2); ep['MaxPool_4a_3x3'] = net
    for name in ['Mixed_4b', 'Mixed_4c', 'Mixed_4d', 'Mixed_4e', 'Mixed_4f']:
        net = _mixed(net, W, name); ep[name] = net
    net = max_pool_same(net, 2, 2); ep['MaxPool_5a_2x2'] = net
    for name in ['Mixed_5b', 'Mixed_5c']:
        net = _mixed(net, W, name); ep[name] = net
    B, H, W_, C = net.shape
    # slim.avg_pool2d(net, [7,7], stride=1) VALID  (inception_v1.py:326)
    Ho, Wo = H - 7 + 1, W_ - 7 + 1
    pooled = np.zeros((B, Ho, Wo, C), dtype=net.dtype)
    for i in range(Ho):
        for j in range(Wo):
            pooled[:, i, j, :] = net[:, i:i + 7, j:j + 7, :].mean(axis=(1, 2))
    ep['AvgPool_0a_7x7'] = pooled
    return pooled, ep


def layer_norm(x, gamma, beta, eps=1e-12):
    """tf.contrib.layers.layer_norm over the last axis: tf.nn.moments
    (biased variance) + tf.nn.batch_normalization(variance_epsilon=1e-12):
    inv = rsqrt(var+eps)*gamma ; x*inv + (beta - mean*inv)."""
    dt = x.dtype
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    inv = (1.0 / np.sqrt(var + dt.type(eps))) * gamma.astype(dt)
    return x * inv + (beta.astype(dt) - mean * inv)


def encoder(images, W, config):
    """src/model_base.py:56-104 -> (im_embed [B,1024], cnn_fmaps [B,196,C])."""
    net, ep = inception_v1(images, W)
    im_embed = net[:, 0, 0, :]                                 # squeeze :93
    if config.legacy:                                          # :80-91
        ENC = 'Model/encoder/'
        im_embed = np.tanh(layer_norm(im_embed, W[ENC + 'LN_tanh/gamma'], W[ENC + 'LN_tanh/beta']))
        im_embed = im_embed @ W[ENC + 'im_embed/weight'].astype(im_embed.dtype)
    fm = ep[config.cnn_fm_attention]                           # :98-103
    B, H, W_, C = fm.shape
    return im_embed, fm.reshape(B, H * W_, C), ep


def preprocess_eval(images_u8, out_hw=(224, 224), resize=256):
    """common/inputs/preprocessing/inception_preprocessing_radix.py:270-273 + :229-235 (is_training=False):
    tf.image.convert_image_dtype(uint8 -> float32) = x * (1/255); tf.image.resize_bilinear to 256 x 256 with TF
    r1.9's align_corners=False rule (in = out * in_size / out_size; top = floor, bottom = min(top + 1, size - 1);
    lerp along x first, then y -- resize_bilinear_op.cc); resize_image_with_crop_or_pad (central crop offset
    (256 - out) // 2, zero pad offset (out - 256) // 2); (x - 0.5) * 2.   PARITY UNPINNED against TF (restated)."""
    x = np.asarray(images_u8)
    assert x.dtype == np.uint8 and x.ndim == 4 and x.shape[3] == 3
    B, H, W, _ = x.shape
    f = x.astype(np.float32) * np.float32(1.0 / 255.0)
    sy, sx = np.float32(H) / np.float32(resize), np.float32(W) / np.float32(resize)
    fy = np.arange(resize, dtype=np.float32) * sy
    fx = np.arange(resize, dtype=np.float32) * sx
    y0 = np.floor(fy).astype(np.int64); x0 = np.floor(fx).astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1); x1 = np.minimum(x0 + 1, W - 1)
    ly = (fy - y0.astype(np.float32))[None, :, None, None]
    lx = (fx - x0.astype(np.float32))[None, None, :, None]
    tl = f[:, y0][:, :, x0]; tr = f[:, y0][:, :, x1]
    bl = f[:, y1][:, :, x0]; br = f[:, y1][:, :, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    r = top + (bot - top) * ly                                        # [B, 256, 256, 3]
    oh, ow = out_hw
    out = np.zeros((B, oh, ow, 3), np.float32)
    ys = (resize - oh) // 2 if oh <= resize else 0; yd = 0 if oh <= resize else (oh - resize) // 2
    xs = (resize - ow) // 2 if ow <= resize else 0; xd = 0 if ow <= resize else (ow - resize) // 2
    hh, ww = min(oh, resize), min(ow, resize)
    out[:, yd:yd + hh, xd:xd + ww] = r[:, ys:ys + hh, xs:xs + ww]
    return (out - np.float32(0.5)) * np.float32(2.0)


def preprocess_train(images_u8, crop_yx, flip=None, out_hw=(224, 224), resize=256):
    """common/inputs/preprocessing/inception_preprocessing_radix.py:270-273 + :158-201 (is_training=True) with the random
    draws given: convert_image_dtype, resize_bilinear(256, 256), tf.image.random_flip_left_right (flip[b]),
    tf.random_crop at crop_yx[b], (x - 0.5) * 2.   PARITY UNPINNED against TF (restated)."""
    x = np.asarray(images_u8)
    B = x.shape[0]
    full = preprocess_eval(x, (resize, resize), resize)                 # resized image, already standardised
    out = np.empty((B, out_hw[0], out_hw[1], 3), np.float32)
    for b in range(B):
        im = full[b, :, ::-1] if (flip is not None and flip[b]) else full[b]
        y0, x0 = int(crop_yx[b][0]), int(crop_yx[b][1])
        out[b] = im[y0:y0 + out_hw[0], x0:x0 + out_hw[1]]
    return out

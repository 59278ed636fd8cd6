"""Generate tests/golden/lena_psnr.npz: BASELINE config 4 (main_lena.py, Lena denoise, K = 5/10/15 layers) evaluated by
the UNMODIFIED reference class.  Test infrastructure; run in the build container only (needs /root/reference):

    python oracle/make_golden_lena.py

The reference's data set `lena_pepper_01.mat` (dictionary D, noisy/clean patches) and its checkpoint are missing from the
repository, so the fixture is a stand-in built from what IS there:
  * gt_x: reference/lena_01.jpg (512x512 gray) cut into 1024 16x16 patches in the inverse order of trans2image
    (main_lena.py:106-115), stored as uint8 (the metric works on [0,1] floats = uint8 / 255);
  * test_x = gt_x with 10 % salt-and-pepper noise, D = an overcomplete 2-D DCT dictionary (256 x 512, unit columns), the
    lena parameters (beta1/beta2 (m x bs) = 0.5*(1 + 0.1*randn), fc weights = (0.9/||D^T D||_2)*(D^T + 1e-3*randn), i.e.
    an UNTRAINED classical-LADMM-like operating point: PSNR rises with depth but stays far from a trained model's) --
    all drawn from torch CPU generators whose seed is in the fixture, so the test rebuilds them bit for bit
    (`lena_case` below is imported by the tests);
  * psnr/K: the evaluation loop of main_lena.py:240-262 run on the reference class: mean over the 51 batches of 20
    patches of mse(255*gt, 255*D@Z_k), PSNR_k = -10*log10(mse_k) + 48.131, for every layer k.
"PSNR parity" is therefore module-vs-reference on identical data and weights, not the README's absolute 35.61 dB.
"""
import os
import sys

import numpy as np
import torch

M, D, BS, N_TEST = 256, 512, 20, 1024
DEPTHS = (5, 10, 15)
SEED = 1126


def patches_from_image(img):
    """(512,512) array -> (256, 1024): column `count` is the 16x16 block (ii,jj) transposed and flattened, so that
    trans2image (main_lena.py:106-115) maps it back."""
    out = np.zeros((256, 1024), dtype=img.dtype)
    count = 0
    for ii in range(0, 512, 16):
        for jj in range(0, 512, 16):
            out[:, count] = img[ii:ii + 16, jj:jj + 16].T.reshape(256)
            count += 1
    return out


def dct_dictionary():
    """Overcomplete separable DCT for 16x16 patches: kron(16x16 DCT, 16x32 overcomplete DCT), unit columns (float64 math,
    rounded once to float32)."""
    def dct(n, k):
        Dm = np.zeros((n, k))
        for j in range(k):
            v = np.cos(np.arange(n) * j * np.pi / k)
            if j > 0:
                v = v - v.mean()
            Dm[:, j] = v / np.linalg.norm(v)
        return Dm
    A = np.kron(dct(16, 16), dct(16, 32))
    A = A / np.sqrt((A * A).sum(axis=0, keepdims=True))
    return torch.from_numpy(A.astype(np.float32))


def lena_case(gt_u8, K, seed=SEED):
    """Everything the evaluation needs, rebuilt deterministically from the stored patches and a seed."""
    g = torch.Generator().manual_seed(seed + K)
    gt = torch.from_numpy(gt_u8.astype(np.float32) / 255.0)
    r = torch.rand(gt.shape, generator=g)
    noisy = gt.clone()
    noisy[r < 0.05] = 0.0                                   # pepper
    noisy[(r >= 0.05) & (r < 0.10)] = 1.0                   # salt
    A = dct_dictionary()
    step = 0.9 / torch.linalg.matrix_norm(A.double().t() @ A.double(), ord=2).item()
    Z0 = torch.rand(D, BS, generator=g) / D                 # main_lena.py:174
    E0 = torch.zeros(M, BS); L0 = torch.zeros(M, BS)
    sd = {}
    for k in range(K):                                      # registration order of main_lena.py:30-37
        sd["beta1.%d" % k] = 0.5 * (1.0 + 0.1 * torch.randn(M, BS, generator=g))
    for k in range(K):
        sd["beta2.%d" % k] = 0.5 * (1.0 + 0.1 * torch.randn(M, BS, generator=g))
    for k in range(K):
        sd["fc.%d.weight" % k] = (step * (A.t() + 1e-3 * torch.randn(D, M, generator=g))).float().contiguous()
    return dict(gt=gt, noisy=noisy, A=A, Z0=Z0, E0=E0, L0=L0, sd=sd)


def psnr_per_layer(forward, case, K):
    """main_lena.py:240-262 with `forward(x) -> (Z, E, L)`."""
    nb = N_TEST // BS
    mse = torch.zeros(K, dtype=torch.float64)
    for j in range(nb):
        x = case["noisy"][:, j * BS:(j + 1) * BS].contiguous()
        gt = case["gt"][:, j * BS:(j + 1) * BS].contiguous()
        Z = forward(x)[0]
        for k in range(K):
            rec = case["A"].to(Z[k].device).mm(Z[k])
            mse[k] += torch.nn.functional.mse_loss(255 * gt.to(rec.device), 255 * rec).item()
    mse /= nb
    return -10 * torch.log10(mse) + 48.131


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import load_reference as lr
    from PIL import Image
    if not lr.reference_available():
        sys.exit("reference not found at %s" % lr.REFERENCE_ROOT)
    img = np.array(Image.open(os.path.join(lr.REFERENCE_ROOT, "lena_01.jpg")).convert("L"), dtype=np.uint8)
    assert img.shape == (512, 512)
    gt_u8 = patches_from_image(img)
    blob = dict(gt_u8=gt_u8, seed=np.array(SEED), depths=np.array(DEPTHS))
    for K in DEPTHS:
        case = lena_case(gt_u8, K)
        ref = lr.build("lena", M, 10000, D, BS, case["A"], case["Z0"], case["E0"], case["L0"], K)
        ref.load_state_dict(case["sd"])
        with torch.no_grad(), lr.cuda_is_identity():
            psnr = psnr_per_layer(lambda x: ref(x), case, K)
        blob["psnr/%d" % K] = psnr.numpy()
        blob["noisy_sum/%d" % K] = np.array(case["noisy"].double().sum().item())
        print("K=%d  PSNR per layer:" % K, np.round(psnr.numpy(), 3))
    noisy_psnr = -10 * np.log10(((255 * (case["noisy"] - case["gt"])[:, :1020]) ** 2).mean().item()) + 48.131
    print("PSNR of the noisy input: %.3f dB" % noisy_psnr)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lena_psnr.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")

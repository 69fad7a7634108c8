"""Oracle (test infrastructure): TI-SV EER threshold sweep.

Follows /root/reference/train_speech_embedder.py:127-149 operation for operation, keeping
its arithmetic types: thresholds are Python doubles ``0.01*i+0.5`` that torch rounds to
float32 for the ``>`` compare (:134-135); per-speaker counts are float32 sums (:137-141);
the Python ``sum`` adds them sequentially in float32; the divisions by (N-1), M/2 and N
are sequential float32 divides; selection is the strict ``diff > |FAR-FRR|`` from
``diff = 1`` (:132,144), so the first minimum wins and EER stays 0 when nothing beats 1.
"""
import numpy as np

THRESHOLDS = [0.01 * i + 0.5 for i in range(50)]        # train_speech_embedder.py:134


def eer_sweep(sim, thresholds=None):
    """sim (N, Mv, N) float32 -> (EER, thresh, FAR, FRR); EER/FAR/FRR float32, thresh double."""
    sim = np.asarray(sim, dtype=np.float32)
    N, Mv, _ = sim.shape
    f32 = np.float32
    thresholds = THRESHOLDS if thresholds is None else thresholds
    idx = np.arange(N)
    diff = 1
    EER = 0
    EER_thresh = 0
    EER_FAR = 0
    EER_FRR = 0
    for thres in thresholds:
        above = sim > f32(thres)
        cnt_all = above.reshape(N, -1).sum(axis=1).astype(np.float32)   # exact: counts < 2^24
        cnt_diag = above[idx, :, idx].sum(axis=1).astype(np.float32)
        far_sum = f32(0)
        frr_sum = f32(0)
        for i in range(N):                                              # Python sum(): sequential fp32
            far_sum = f32(far_sum + f32(cnt_all[i] - cnt_diag[i]))
            frr_sum = f32(frr_sum + f32(f32(Mv) - cnt_diag[i]))
        FAR = f32(f32(f32(far_sum / f32(N - 1.0)) / f32(float(Mv))) / f32(N))
        FRR = f32(f32(frr_sum / f32(float(Mv))) / f32(N))
        d = f32(abs(f32(FAR - FRR)))
        if diff > d:
            diff = d
            EER = f32(f32(FAR + FRR) / f32(2))
            EER_thresh = thres
            EER_FAR = FAR
            EER_FRR = FRR
    return EER, EER_thresh, EER_FAR, EER_FRR


def eer_counts(sim, thresholds=None):
    """Integer intermediate of the sweep: (50, N) counts above threshold, all and diagonal."""
    sim = np.asarray(sim, dtype=np.float32)
    N = sim.shape[0]
    thresholds = THRESHOLDS if thresholds is None else thresholds
    idx = np.arange(N)
    call = np.zeros((len(thresholds), N), dtype=np.int64)
    cdiag = np.zeros((len(thresholds), N), dtype=np.int64)
    for t, thres in enumerate(thresholds):
        above = sim > np.float32(thres)
        call[t] = above.reshape(N, -1).sum(axis=1)
        cdiag[t] = above[idx, :, idx].sum(axis=1)
    return call, cdiag

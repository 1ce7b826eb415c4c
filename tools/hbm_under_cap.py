"""What does a plain device copy reach while the GPU sits at its power cap?

MEASURED_PEAKS.json's 6542 GB/s is a copy timed ALONE on a cool GPU.  Inside the 23 ms step every HBM-bound kernel runs
at 5.0-5.2 TB/s and the SM clock sits at 1.39-1.6 GHz under the 1 kW cap.  This tool times the same copy (a) alone and
(b) interleaved with bf16 matmuls that hold the GPU at the cap (the copy launches sit between two matmuls on the same
stream, CUDA events around each copy), and samples the SM clock meanwhile.  If (b) stays at the alone figure the
in-step kernels lose to latency (more bytes in flight would help); if it drops with the clock, the in-step HBM roofline
itself is lower than the alone figure and those kernels already sit on it.

    python tools/hbm_under_cap.py > gpurun_out/hbm_under_cap.json
"""
import json
import statistics
import threading
import time

import torch


def _clock_sampler(stop, out):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.02)
    except Exception as e:                                               # noqa: BLE001
        out.append(("nvml unavailable", str(e)))


def timed_copies(dst, src, n, between=None):
    evs = []
    for _ in range(n):
        if between is not None:
            between()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dst.copy_(src)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return ms


def main():
    dev = "cuda:0"
    res = {}
    for label, nbytes in (("2GiB", 2 << 30), ("256MiB", 256 << 20)):
        src = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=dev).normal_()
        dst = torch.empty_like(src)
        A = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        Bm = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        C = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize()
        alone = timed_copies(dst, src, 10)
        gbs = lambda ms: 2 * nbytes / ms / 1e6                           # noqa: E731
        stop, samples = threading.Event(), []
        th = threading.Thread(target=_clock_sampler, args=(stop, samples))

        def burn(k=4):
            for _ in range(k):
                torch.matmul(A, Bm, out=C)

        for _ in range(300):                                             # ~0.25 s: reach the cap before measuring
            torch.matmul(A, Bm, out=C)
        torch.cuda.synchronize()
        th.start()
        capped = timed_copies(dst, src, 60, between=burn)
        stop.set()
        th.join()
        good = [s for s in samples if isinstance(s[0], int)]
        res[label] = {
            "copy_alone_gbs_best": round(gbs(min(alone)), 1), "copy_alone_gbs_median": round(gbs(statistics.median(alone)), 1),
            "copy_between_matmuls_gbs_best": round(gbs(min(capped[10:])), 1),
            "copy_between_matmuls_gbs_median": round(gbs(statistics.median(capped[10:])), 1),
            "sm_mhz_median_under_load": statistics.median([s[0] for s in good]) if good else None,
            "mem_mhz_median_under_load": statistics.median([s[1] for s in good]) if good else None,
            "power_w_median": round(statistics.median([s[2] for s in good]), 1) if good else None,
        }
        del src, dst
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""whole-encode cases with committed golden md5s (tests/golden/reference_ivf.json, made by tests/golden/make_golden.py
from the reference encoder on the CPU; checked on the GPU by tests/test_gpu_e2e.py::test_ivf_matches_the_committed_reference_md5)"""
IVF_CASES = {
    # name: (w, h, frames, host args) -- the first four are the cases of tests/test_gpu_e2e.py
    "qcif": (176, 144, 14, ["-qmin", 20, "-qmax", 44, "-g", 12, "-altref-range", 4, "-partitions", 2, "-threads", 2]),
    "cif": (352, 288, 20, ["-qmin", 24, "-qmax", 24, "-g", 60, "-altref-range", 5, "-partitions", 1, "-threads", 2]),
    "cif_ssim": (352, 288, 12, ["-qmin", 10, "-qmax", 50, "-g", 30, "-altref-range", 3, "-partitions", 4, "-threads", 12,
                                "-SSIM-target", "93"]),
    "odd": (200, 120, 8, ["-qmin", 30, "-qmax", 30, "-g", 8, "-altref-range", 2, "-partitions", 8, "-threads", 12]),
    "cif_60_baseline": (352, 288, 60, ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
    "1080p_12": (1920, 1080, 12, ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
    "2160p_6": (3840, 2160, 6, ["-qmin", 24, "-qmax", 24, "-g", 150, "-altref-range", 5, "-partitions", 8, "-threads", 12]),
}

"""dafne_b200 -- B200-native (sm_100a) inference hot path for the DAFNe oriented-object detector.

Only the batched-inference path exists here (dense forward + rotated-box post-processing), behind the reference's
own Python surface (`OneStageDetector(cfg)(batched_inputs) -> [{"instances": Instances}]`). All arithmetic runs in
hand-written CUDA kernels reached through the C ABI of ``libdafne_b200.so``; there is no CPU or eager fallback.
"""
__version__ = "0.1.0"

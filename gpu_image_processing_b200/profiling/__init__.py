from .ncu_profiler import check_ncu_available, get_common_ncu_metrics, profile_kernel_with_ncu  # noqa: F401

"""Drop-in replacements for /root/reference/repellency/repellency_methods_{fast,fast_sdv3,threshold}.py.

    from safe_denoiser_b200.repellency.repellency_methods_threshold import get_repellency_method
"""

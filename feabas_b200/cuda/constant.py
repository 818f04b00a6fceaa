"""Constants shared with the reference (feabas/constant.py:39-41)."""
FFT_CONF_NONE = 0
FFT_CONF_STD = 1
FFT_CONF_MIRROR = 2

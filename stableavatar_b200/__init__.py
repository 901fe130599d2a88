"""B200-native (sm_100a) denoising hot path of StableAvatar: C-ABI kernels (csrc/) + host-side mirrors of the reference classes."""

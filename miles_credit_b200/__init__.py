"""B200-native forecast forward step for NCAR CREDIT's WXFormer/CrossFormer."""

"""B200 mirror of the reference's ``quantization_utils`` package (quant_utils + quant_modules)."""

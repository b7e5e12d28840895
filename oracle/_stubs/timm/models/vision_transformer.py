class VisionTransformer:  # placeholder for the import at deit_models.py:11
    pass

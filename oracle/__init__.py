"""TEST INFRASTRUCTURE: CPU oracle of the deformer hot path. Not part of the product."""

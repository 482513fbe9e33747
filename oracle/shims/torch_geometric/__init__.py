"""Shim package for torch_geometric (absent from this image). TEST INFRASTRUCTURE ONLY."""

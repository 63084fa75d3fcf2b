"""dask stub (test infrastructure only) - see oracle/_stubs/xarray/__init__.py."""

"""Drop-in replacement for the reference's `models` package (models/layers.py, models/pangu_model.py):
same class names, constructor / forward signatures and the same 223 state_dict keys, executed by the
sm_100a kernels of libpangu_b200.so.  Put this directory's parent (`pangu-pytorch-demo_b200/`) on
sys.path where the reference scripts put their repo root."""

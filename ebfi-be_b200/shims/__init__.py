"""Drop-in replacements for the reference's two compiled extension modules.

`_ext`               <- models/DCNv2/src/vision.cpp:4-9      (imported by models/DCNv2/dcn_v2.py:13)
`kernelconv2d_cuda`  <- models/FAC/kernelconv2d/KernelConv2D_cuda.cpp:58-61
                                                         (imported by .../KernelConv2D.py:8)

Put this directory on sys.path (or call ebfi_be_b200.install_shims()) and the reference's
unmodified Python wrappers run on the sm_100a kernels.
"""

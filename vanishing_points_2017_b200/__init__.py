"""B200-native hot path of fkluger/vanishing_points_2017.

The EM stage runs up to 8 groups of images on their own CUDA streams next to the
stream of the sphere-mapping / CNN stages.  With the driver's default of 8
hardware work queues the streams alias onto the same queues and falsely
serialise (measured: 5 or more groups cost +30 % on the EM stage), so the
package asks for 32 queues.  The variable is read when the CUDA context is
created: import this package before the first CUDA call of the process, or
export CUDA_DEVICE_MAX_CONNECTIONS yourself.
"""
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

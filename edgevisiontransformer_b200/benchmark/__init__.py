"""B200 backend, the sibling of the reference's benchmark/{tensorrt,openvino} and utils.trt_benchmark."""
from .b200 import b200_benchmark, build_model, main  # noqa: F401

"""Introspection of captured CUDA graphs through the driver API (libcuda, ctypes): how many kernel nodes a replay launches.
bench.py reports this as `gpu_launches` -- a count of what the graph holds, not an estimate."""
import ctypes

_KERNEL = 0      # CU_GRAPH_NODE_TYPE_KERNEL
_cuda = None


def _drv():
    global _cuda
    if _cuda is None:
        _cuda = ctypes.CDLL("libcuda.so.1")
    return _cuda


def count_nodes(graph):
    """torch.cuda.CUDAGraph (captured with keep_graph=True) -> {node type: count}"""
    drv = _drv()
    g = ctypes.c_void_p(graph.raw_cuda_graph())
    n = ctypes.c_size_t(0)
    if drv.cuGraphGetNodes(g, None, ctypes.byref(n)) != 0:
        raise RuntimeError("cuGraphGetNodes failed")
    nodes = (ctypes.c_void_p * n.value)()
    if drv.cuGraphGetNodes(g, nodes, ctypes.byref(n)) != 0:
        raise RuntimeError("cuGraphGetNodes failed")
    out = {}
    for i in range(n.value):
        t = ctypes.c_int(0)
        if drv.cuGraphNodeGetType(ctypes.c_void_p(nodes[i]), ctypes.byref(t)) != 0:
            raise RuntimeError("cuGraphNodeGetType failed")
        out[t.value] = out.get(t.value, 0) + 1
    return out


def count_kernel_nodes(graph):
    return count_nodes(graph).get(_KERNEL, 0)

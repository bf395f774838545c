from .wavenet import WaveNet

__all__ = ["WaveNet", "receptive_field_size"]


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    """Receptive field in samples of a dilated stack (nnsvs/wavenet/__init__.py:6-33)."""
    assert total_layers % num_cycles == 0
    per_cycle = total_layers // num_cycles
    return (kernel_size - 1) * sum(dilation(i % per_cycle) for i in range(total_layers)) + 1

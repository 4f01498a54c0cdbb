"""Entry point mirroring postprocessing/decoding.py:6-53."""
import torch


def decode(model, input, data_loader=None, expected=None, method="greedy", beam_width=3):
    """Same arguments, return value and error behaviour as the reference's
    ``decode``: greedy -> int64 [B, expected.size(1)-1] on the input's device
    (logits -> transpose -> topk(1) over the vocab axis, decoding.py:35-40);
    beam -> int64 [B, expected.size(-1)-1] on the CPU (decoding.py:42-48)."""
    if method == "greedy":
        output = model(input=input, expected=expected, is_train=False, teacher_forcing_ratio=0.0)
        decoded_values = output.transpose(1, 2)
        _, sequence = torch.topk(decoded_values, 1, dim=1)
        sequence = sequence.squeeze(1)
    elif method == "beam":
        sequence = model.beam_search(input=input, data_loader=data_loader, beam_width=beam_width,
                                     max_sequence=expected.size(-1) - 1)
    else:
        raise NotImplementedError(f"There's no '{method}' type yet.")
    return sequence

"""The per-batch step of the reference's training / validation loop (``libs/trainer.py:165-196``,
``Trainer.inference_one_batch``) over the CUDA forward and the device loss.

``phase='val'`` is the reference's validation step verbatim: inputs to the device, ``model.eval()``, forward and
``FuseLoss`` under ``torch.no_grad()``, every ``*loss*`` stat converted to a python float.  ``phase='train'`` needs the backward
of every CUDA stage (not built): it raises instead of running a step that cannot update the weights.
"""
import torch


def inference_one_batch(model, loss, input_dict, phase, device="cuda"):
    assert phase in ["train", "val"]
    for key, value in input_dict.items():  # libs/trainer.py:169-171
        if not isinstance(value, list):
            input_dict[key] = value.to(device)
    if phase == "train":
        raise NotImplementedError("training step: MotionNet's CUDA stages have no backward pass (loss gradients stop at the "
                                  "network outputs); use phase='val' or the reference model for training")
    model.eval()
    with torch.no_grad():
        predictions = model(input_dict)
        stats = loss(predictions, input_dict)
    for key, value in stats.items():  # libs/trainer.py:190-193
        if key.find("loss") != -1:
            stats[key] = float(value.detach()) if isinstance(value, torch.Tensor) else float(value)
    return stats

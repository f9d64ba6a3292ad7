"""Host-side helpers of the classification hot path with the reference's names
(SimpleAICV/classification/common.py:668-685 AverageMeter, :688-709 AccMeter, :843-881 get_amp_type).
Transforms / collaters / datasets are CPU data code and out of scope."""
import torch


class AverageMeter:
    """Running average (sum of val*n over count)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class AccMeter:
    """Accumulates top-1 / top-5 correct counts; ``compute()`` fills ``acc1`` / ``acc5`` as fractions."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.acc1_correct_num = 0
        self.acc5_correct_num = 0
        self.sample_num = 0
        self.acc1 = 0
        self.acc5 = 0

    def update(self, acc1_correct_num, acc5_correct_num, sample_num):
        self.acc1_correct_num += acc1_correct_num
        self.acc5_correct_num += acc5_correct_num
        self.sample_num += sample_num

    def compute(self):
        self.acc1 = float(self.acc1_correct_num) / self.sample_num if self.sample_num != 0 else 0
        self.acc5 = float(self.acc5_correct_num) / self.sample_num if self.sample_num != 0 else 0


def get_amp_type(model):
    """The autocast dtype the reference's loop would select for the device ``model`` lives on.

    Same decision procedure as the reference (compute capability >= 8, torch >= 1.10, CUDA >= 11, device
    name in a whitelist) with 'B200' (and 'B300' / 'GB200') added to the whitelist: the unmodified list
    lacks Blackwell parts, so the reference falls back to fp16 + GradScaler on a B200.
    The kernels of this package always compute in bf16 with fp32 accumulation; the value is what
    ``train_config`` / the loops report and what a torch-side tail (criterion) may autocast to."""
    device = next(model.parameters()).device
    if device.type != 'cuda':
        return torch.bfloat16
    properties = torch.cuda.get_device_properties(device)
    ok = properties.major >= 8
    ver = tuple(int(v) for v in torch.__version__.split('+')[0].split('.')[:2])
    ok = ok and ver >= (1, 10)
    if torch.version.cuda and float(torch.version.cuda.split('.')[0]) < 11:
        ok = False
    name = torch.cuda.get_device_name(device)
    whitelist = ['RTX PRO 6000', 'H20', 'L20', 'L40', '4090', '5090', 'A100', 'A800', 'H100', 'H800',
                 'B200', 'B300', 'GB200']
    ok = ok and any(n in name for n in whitelist)
    return torch.bfloat16 if ok else torch.float16

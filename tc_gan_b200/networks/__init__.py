"""
Callers of the SSN hot path with the reference's own names: the BPTT tuning-curve
generator (tc_gan/networks/ssn.py), the WGAN and conditional-WGAN learners (networks/wgan.py, cwgan.py), the
truth-dataset provider (networks/dataset.py) and the grid helper (networks/utils.py).  Theano graphs become
torch autograd over the CUDA operators of ``tc_gan_b200.torch_ops``; Lasagne's MLP critic becomes torch.nn.
"""

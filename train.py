#!/usr/bin/env python3
"""Social Ways trajectory prediction -- B200-native drop-in for the reference train.py.

Same command line as the reference (train.py:19-50: --batch-size --epochs --model --latent-dim
--d-learning-rate --g-learning-rate --unrolling-steps --hidden-size --dataset), same defaults, same
printed lines, same checkpoint dictionary (train.py:653-663) and prediction dumps (train.py:591-599).
Flags ADDED here (the reference hard-codes these as module constants / paths, train.py:56-57,83):
  --use-social      flip the `use_social` constant (reference default False, train.py:83)
  --input-file      dataset npz (reference: '../hotel-8-12.npz')
  --model-file      checkpoint path (reference: '../trained_models/<model>-<dataset>.pt')
  --out-dir         where test() dumps go (reference: '../medium/<dataset>/<model>/<epoch>')
  --seed            seeds numpy and torch (the reference seeds nothing; under torchrun rank 0's seed is broadcast)
  --test-samples    K of the periodic test() call (reference: 128, train.py:668)
  --cuda-graph      replay the whole GAN iteration from a CUDA graph per batch shape (5x at batch 256)
  --fused-adam      one kernel per optimiser step on flat buffers (fused_optim.FlatAdam); under torchrun it also carries
                    the gradient all-reduce of the sharded step over NVLink peer memory, which is what lets --cuda-graph
                    capture the multi-GPU iteration
  --native          the iteration as ~30 launches of this library's own kernels replayed from one CUDA graph per mini-batch
                    shape (no autograd / cuBLAS / ATen; implies --fused-adam): 0.4 ms per iteration at batch 256 on the toy set
  --device-noise S  (with --native) draw the per-iteration latent noise on the GPU (Philox, seed S) instead of torch.rand
                    on the CPU: same distribution, different stream -- the host generator otherwise bounds large batches
--hidden-size: the sm_100a kernels are built for the default 64 (every BASELINE configuration); other values raise.
Under `torchrun --nproc-per-node N train.py ...` the scenes of every mini-batch are sharded over the N GPUs.
All arithmetic runs in the sm_100a kernels of socialways_b200 (no CPU fallback).
"""
import argparse
import os

import numpy as np
import torch
from tqdm import trange

parser = argparse.ArgumentParser(description='Social Ways trajectory prediction.')
parser.add_argument('--batch-size', '--b', type=int, default=256, metavar='N',
                    help='input batch size for training (default: 256)')
parser.add_argument('--epochs', '--e', type=int, default=1000, metavar='N',
                    help='number of epochs to train (default: 1000)')
parser.add_argument('--model', '--m', default='socialWays', choices=['socialWays'],
                    help='pick a specific network to train (default: "socialWays")')
parser.add_argument('--latent-dim', '--ld', type=int, default=10, metavar='N',
                    help='dimension of latent space (default: 10)')
parser.add_argument('--d-learning-rate', '--d-lr', type=float, default=1E-3, metavar='N',
                    help='learning rate of discriminator (default: 1E-3)')
parser.add_argument('--g-learning-rate', '--g-lr', type=float, default=1E-4, metavar='N',
                    help='learning rate of generator (default: 1E-4)')
parser.add_argument('--unrolling-steps', '--unroll', type=int, default=1, metavar='N',
                    help='number of steps to unroll gan (default: 1)')
parser.add_argument('--hidden-size', '--h-size', type=int, default=64, metavar='N',
                    help='size of network intermediate layer (default: 64; the sm_100a kernels are built for 64 only, '
                         'other values raise SocialWaysCudaError)')
parser.add_argument('--dataset', '--data', default='hotel', choices=['hotel'],
                    help='pick a specific dataset (default: "hotel")')
parser.add_argument('--use-social', action='store_true')
parser.add_argument('--input-file', default='../hotel-8-12.npz')
parser.add_argument('--model-file', default=None)
parser.add_argument('--out-dir', default=None)
parser.add_argument('--seed', type=int, default=None)
parser.add_argument('--test-samples', type=int, default=128)
parser.add_argument('--cuda-graph', action='store_true',
                    help='capture each mini-batch shape of train() into a CUDA graph and replay it')
parser.add_argument('--fused-adam', action='store_true',
                    help='flat-buffer Adam in one kernel (with the gradient all-reduce fused in under torchrun)')
parser.add_argument('--native', action='store_true',
                    help='run the iteration on the library\'s own kernels only (native_step.py), one CUDA graph per batch shape')
parser.add_argument('--device-noise', type=int, default=None, metavar='SEED',
                    help='with --native: latent noise drawn on the GPU (Philox) instead of torch.rand on the CPU')


def main():
    args = parser.parse_args()
    from socialways_b200.trainer import SocialWaysTrainer
    model_file = args.model_file or '../trained_models/' + args.model + '-' + args.dataset + '.pt'
    print(os.path.dirname(os.path.realpath(__file__)))
    device = "cuda"
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    seed = args.seed
    if distributed:                                                  # torchrun: one rank per GPU, scenes sharded
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = f"cuda:{local}"
        dist.init_process_group("nccl", device_id=torch.device(device))
        if args.cuda_graph and not (args.fused_adam or args.native):
            raise SystemExit("--cuda-graph under torchrun needs --fused-adam (NCCL calls are not captured)")
        # the sharded step relies on IDENTICAL numpy / torch CPU RNG streams on every rank (labels and noise are drawn for
        # the global mini-batch everywhere): without --seed rank 0 picks one and every rank adopts it
        box = [int(np.random.SeedSequence().entropy % (2 ** 31)) if seed is None else seed]
        dist.broadcast_object_list(box, src=0)
        seed = box[0]
    if seed is not None:
        np.random.seed(seed)
        torch.manual_seed(seed)
    data = np.load(args.input_file)
    tr = SocialWaysTrainer(data, batch_size=args.batch_size, hidden_size=args.hidden_size,
                           use_social=args.use_social, n_unrolling_steps=args.unrolling_steps,
                           lr_g=args.g_learning_rate, lr_d=args.d_learning_rate, cuda_graph=args.cuda_graph and not args.native,
                           fused_adam=args.fused_adam or args.native, device=device)
    if distributed:                                                  # replicas start from rank 0's parameters, whatever built them
        for p in list(tr.generator.parameters()) + list(tr.D.parameters()):
            dist.broadcast(p.data, src=0)
        tr.generator.invalidate_packs()
    print(args.input_file, ' # Training samples: ', tr.n_train_samples)
    print('hidden dim = %d | lr(G) =  %.5f | lr(D) =  %.5f' % (args.hidden_size, args.g_learning_rate, args.d_learning_rate))
    if os.path.isfile(model_file):                                   # train.py:622-637
        print('Loading model from ' + model_file)
        start_epoch = tr.load_state(torch.load(model_file))
    else:
        start_epoch = 1
    for epoch in trange(start_epoch, args.epochs + 1):               # train.py:646-668
        tr.epoch = epoch
        if args.native:
            tr.train_native(device_noise_seed=args.device_noise)
        else:
            (tr.train_graphed if args.cuda_graph else tr.train)()
        if tr.rank == 0:                                            # replicas are bit-identical: rank 0 writes files
            if epoch % 50 == 0:
                print('Saving model to file ...', model_file)
                os.makedirs(os.path.dirname(os.path.abspath(model_file)), exist_ok=True)
                torch.save(tr.state(), model_file)
            if epoch % 5 == 0:
                wr_dir = os.path.join(args.out_dir or ('../medium/' + args.dataset + '/' + args.model), str(epoch))
                os.makedirs(wr_dir, exist_ok=True)
                tr.test(args.test_samples, write_to_file=wr_dir, just_one=True)
        elif epoch % 5 == 0:
            tr.skip_test_rng(args.test_samples, write_to_file=True, just_one=True)   # same CPU noise stream as rank 0
        if distributed and (epoch % 5 == 0 or epoch % 50 == 0):
            # nobody starts the next epoch (and spins on rank 0's flags inside the fused all-reduce kernel) while rank 0
            # is still writing dumps or the checkpoint
            dist.barrier()


if __name__ == '__main__':
    main()

"""Drop-in for the reference's gp_classification.py: variational GP classification of the summed mask labels over pixel
coordinates (KISS-GP run through a Bernoulli likelihood).

Same entry points, module constants and on-disk contract as the reference (file:line are the reference's):
  dataset / mode / n                                  :25-38    module-level switches
  load_images_from_folder(folder)                     :41-50    ./masks/mask_{i}_{label}.png, label parsed from the file name
  prepare_training_data()                             :52-136   H[p] = sum of labels of the masks whose pixel p is 255; train_x = covered
                                                                pixel coordinates (i, j), train_y = H there
  GPClassificationModel()                             :139-156  GridInducingVariationalGP(grid_size=10, grid_bounds=[(0,n),(0,n)]),
                                                                near-zero ConstantMean, RBF x exp(log_outputscale)
  train(train_x, train_y, model, likelihood)          :160-217  30 Adam(lr=0.1) steps on -VariationalMarginalLogLikelihood, checkpoint
  eval_superpixels(model, likelihood)                 :219-264  likelihood(model(x)).mean() at all n x n pixels
  plot_result(predictions)                            :267-330  heat maps (written to ./weighted_mask/ instead of cv2.imshow)

What changed underneath: the O(N * n^2) Python dictionary loop is one HBM-bound kernel (`nib_heatmap_pixels`); the part of
every optimiser step that scales with the n training pixels - expected log-likelihood and its gradients - and the
prediction run in libnib.so kernels (`nib_vgp_loglik_grad`, `nib_vgp_predict`, csrc/ski.cu); the G x G (G = 100) KL algebra
and Adam stay on the host.  gpytorch is not needed (and absent here, pre-0.1 API, unpinned by the reference -> parity
unpinned: the model is restated from its published construction, expectations by Gauss-Hermite quadrature where the
reference's BernoulliLikelihood samples).  The checkpoint is a dict of the model's parameters.
"""
from __future__ import annotations

import os

import cv2
import numpy as np
import torch

from gp_regression import _to_u8, summed_label_heatmap
from network_interpretation_imagenet_b200.vgp import GridVariationalGPClassifier

dataset = 'IMAGENET'
# dataset = 'MNIST'

mode = 'Train'
# mode = 'Eval'

if dataset == 'MNIST':
    n = 28
elif dataset == 'CIFAR':
    n = 32
elif dataset == 'IMAGENET':
    n = 224
else:
    raise Exception("This dataset Not implemented yet")

CHECKPOINT = './gp_saved_checkpoints/imgenet100_epoch10_gp_cls_checkpoint.pth.tar'


def load_images_from_folder(folder):
    img_filenames = []
    labels = []
    for filename in os.listdir(folder):
        label = filename.split('_')[2].split('.')[0]
        img_filenames.append(os.path.join(folder, filename))
        labels.append(label)
    return img_filenames, labels


def prepare_training_data():
    import gp_regression
    gp_regression.n = n                                # the heat-map helper reads the module-level image size
    heat, covered = summed_label_heatmap('./masks/')
    idx = torch.nonzero(covered)                       # row-major (i, j) order, as the reference's nested loops
    train_x = idx.to(torch.float32)
    train_y = heat[covered]
    os.makedirs('./weighted_mask', exist_ok=True)
    cv2.imwrite('./weighted_mask/weighted_mask_heatmap.png', cv2.applyColorMap(_to_u8(heat.cpu().numpy()), cv2.COLORMAP_JET))
    print("train_x.shape: ", train_x.shape)
    print("train_y.shape: ", train_y.shape)
    return train_x, train_y


class BernoulliLikelihood:
    """Stand-in for gpytorch.likelihoods.BernoulliLikelihood (no parameters, :173)."""

    def cuda(self):
        return self

    def train(self):
        return self

    def eval(self):
        return self


class GPClassificationModel(GridVariationalGPClassifier):
    """The reference's model (:139-156): grid_size = 10 over [0, n]^2, constant mean bounded to +-1e-5, log_lengthscale and
    log_outputscale bounded to (-5, 6), all initial values 0."""

    def __init__(self):
        super().__init__(grid_size=10, grid_bounds=((0.0, float(n)), (0.0, float(n))), const_mean_bounds=(-1e-5, 1e-5),
                         log_lengthscale_bounds=(-5.0, 6.0), log_outputscale_bounds=(-5.0, 6.0))


def train(train_x, train_y, model, likelihood):
    model.train()
    likelihood.train()
    num_training_iterations = 30
    model.fit(train_x, train_y, num_training_iterations=num_training_iterations, lr=0.1)     # prints the reference's Iter line
    os.makedirs(os.path.dirname(CHECKPOINT), exist_ok=True)
    torch.save(model.state_dict(), CHECKPOINT)


def eval_superpixels(model, likelihood):
    model_dir = CHECKPOINT
    model.load_state_dict(torch.load(model_dir, weights_only=False))
    model.eval()
    likelihood.eval()
    ii, jj = torch.meshgrid(torch.arange(n, dtype=torch.float32), torch.arange(n, dtype=torch.float32), indexing="ij")
    test_x = torch.stack([ii, jj], -1).reshape(-1, 2).cuda()
    print("test_x.shape")
    print(test_x.shape)
    # the reference walks the n*n pixels in batches of 896 (:241-253); one launch covers them all here
    full_predictions = model.predict_proba(test_x).cpu().numpy()
    print(full_predictions.shape)
    return full_predictions


def plot_result(predictions):
    import gp_regression
    gp_regression.n = n
    heat, _ = summed_label_heatmap('./masks')
    os.makedirs('./weighted_mask', exist_ok=True)
    cv2.imwrite('./weighted_mask/summed_label_training_heatmap.png',
                cv2.applyColorMap(_to_u8(heat.cpu().numpy()), cv2.COLORMAP_JET))
    org_test_gray_img = np.asarray(predictions).reshape(n, n)
    print("org_test_gray_img")
    print(org_test_gray_img)
    cv2.imwrite('./weighted_mask/predicted_class_probability_heatmap.png',
                cv2.applyColorMap(_to_u8(org_test_gray_img), cv2.COLORMAP_JET))


def main():
    model = GPClassificationModel().cuda()
    likelihood = BernoulliLikelihood().cuda()
    if mode == 'Train':
        train_x, train_y = prepare_training_data()
        train(train_x, train_y, model, likelihood)
    elif mode == 'Eval':
        print("start to test the model")
        predictions = eval_superpixels(model, likelihood)
        plot_result(predictions)
    else:
        raise Exception("No such mode")


if __name__ == "__main__":
    main()

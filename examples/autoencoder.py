"""Port of the reference's root autoencoder.py + capsule.py to stock PyTorch (SURVEY.md 8f-4:
the encoders are tiny dense layers that only DRIVE the hot path; the decoder is the ray tracer).

    Autoencoder(scene, n_visible, n_hidden_l1, n_hidden_l2, n_hidden_l3, num_capsule)
        .encoder(X)  tanh -> tanh -> softplus -> one 6-vector (centre, scale) per capsule   autoencoder.py:51-64
        .decoder(robjs) = scene(capsules, robjs)                                            autoencoder.py:66-67
        .cost(X) = sum((X - decoder(encoder(X))[:, :, 0].flatten())**2)                      autoencoder.py:69-75
"""
import numpy as np
import torch

WEIGHT, BIAS = 0, 1


def initialize_weight(n_vis, n_hid, numpy_rng, device):
    """util.py:10-20 ('uniform')"""
    b = np.sqrt(6. / (n_vis + n_hid))
    return torch.tensor(numpy_rng.uniform(low=-b, high=b, size=(n_vis, n_hid)).astype(np.float32), device=device)


class Capsule(object):
    """capsule.py:5-29: a [n_hidden, 6] read-out (3 centre + 3 scale columns) and its bias
    (0, 0, 3, 1/num_caps ...)."""

    def __init__(self, name, n_hidden, n_output, num_caps, device, rng=np.random):
        self.name = name
        bias = np.asarray([0, 0, 3 * num_caps, 1, 1, 1], dtype=np.float32) / num_caps
        hi = 4 * np.sqrt(6. / 6 + n_hidden)               # sic: capsule.py:19 (operator precedence)
        to_center = 0.05 * rng.uniform(low=-hi, high=hi, size=(n_hidden, 3))
        to_radius = 0.0005 * rng.uniform(low=-hi, high=hi, size=(n_hidden, 3))
        self.params = [torch.tensor(np.concatenate((to_center, to_radius), 1).astype(np.float32), device=device),
                       torch.tensor(bias, device=device)]


class Autoencoder(object):
    def __init__(self, scene, n_visible, n_hidden_l1, n_hidden_l2, n_hidden_l3, num_capsule, device='cuda'):
        self.scene = scene
        dev = torch.device(device)
        self.l1_biases = torch.zeros(n_hidden_l1, device=dev)
        self.l2_biases = torch.zeros(n_hidden_l2, device=dev)
        self.l3_biases = torch.zeros(n_hidden_l3, device=dev)
        numpy_rng = np.random.RandomState(1234)
        self.W0 = initialize_weight(n_visible, n_hidden_l1, numpy_rng, dev)
        self.W1 = initialize_weight(n_hidden_l1, n_hidden_l2, numpy_rng, dev)
        self.W2 = initialize_weight(n_hidden_l2, n_hidden_l3, numpy_rng, dev)
        self.params0 = [self.W0, self.W1, self.W2, self.l1_biases, self.l2_biases, self.l3_biases]
        cap_rng = np.random.RandomState(4321)             # the reference uses the unseeded global RNG here
        self.capsules = [Capsule('sphere', n_hidden_l3, 6, num_capsule, dev, cap_rng) for _ in range(num_capsule)]
        self.capsule_params = [p for c in self.capsules for p in c.params]
        self.params = self.params0 + self.capsule_params
        for p in self.params:
            p.requires_grad_(True)

    def encoder(self, X):
        h1 = torch.tanh(X @ self.W0 + self.l1_biases)
        h2 = torch.tanh(h1 @ self.W1 + self.l2_biases)
        h3 = torch.nn.functional.softplus(h2 @ self.W2 + self.l3_biases)
        return [h3 @ c.params[WEIGHT] + c.params[BIAS] for c in self.capsules]

    def decoder(self, robjs):
        return self.scene(self.capsules, robjs)

    def get_reconstruct(self, X):
        return self.decoder(self.encoder(X))

    def cost(self, X):
        recon = self.decoder(self.encoder(X))[:, :, 0].flatten()
        return torch.sum((X - recon) * (X - recon))

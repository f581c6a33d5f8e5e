from .agd import AcceleratedGradientDescent, project_on_nn_cone

__all__ = ["AcceleratedGradientDescent", "project_on_nn_cone"]

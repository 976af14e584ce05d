"""The registry-level drop-in: a pipeline class sduss can put into `EsyMReDPipelineRegistry`
(sduss/model_executor/diffusers/pipelines/__init__.py:26-30) in place of its own.

`make_b200_pipeline(reference_cls, kind)` subclasses the reference's ESyMReD pipeline class (which
only exists where sduss + diffusers are installed, so it is passed in rather than imported), and
overrides exactly the two members that belong to the hot path:

  instantiate_pipeline   wraps the loaded diffusers transformer / unet in the B200 module instead
                         of PatchSD3Transformer2DModel / PatchUNet
                         (pipeline_stable_diffusion_3_esymred.py:24-36, ..._xl_esymred.py:30-41)
  denoising_step         the fused B200 step (same keyword arguments as the reference methods,
                         ..._3_esymred.py:232-388, ..._xl_esymred.py:260-403)

Everything else -- get_sampling_params_cls, __post_init__, prepare_inference (text encoders),
post_inference (VAE), class attributes -- is inherited untouched, so `_ModelRunner`
(sduss/worker/runner/_model_runner.py:109-114,238-264) cannot tell the difference.
With `b200_vae=True` the `vae` sub-module is additionally wrapped in `B200VAEProxy`: the inherited
post_inference is still the reference's code, but its `self.vae.decode(...)` runs on the B200
kernels. With `b200_text_encoders=True` the text encoders are wrapped in
`B200CLIPProxy` / `B200T5Proxy` the same way: the inherited prepare_inference -> encode_prompt is
unchanged and its encoder calls run on the B200 kernels.
"""
from typing import Type

_KINDS = {
    "sd3": ("transformer", "sduss_b200.sd3_transformer", "B200SD3Transformer2DModel",
            "B200StableDiffusion3Pipeline"),
    "sdxl": ("unet", "sduss_b200.unet", "B200UNet", "B200StableDiffusionXLPipeline"),
}


def make_b200_pipeline(reference_cls: Type, kind: str, device: str = "cuda", b200_vae: bool = False,
                       b200_text_encoders: bool = False) -> Type:
    if kind not in _KINDS:
        raise ValueError(f"kind must be one of {sorted(_KINDS)}, got {kind!r}")
    module_key, mod_name, model_name, step_name = _KINDS[kind]

    class B200Pipeline(reference_cls):
        B200_KIND = kind

        @classmethod
        def instantiate_pipeline(cls, **kwargs):
            import importlib
            sub_modules = kwargs.pop("sub_modules", {})
            module = sub_modules.pop(module_key, None)
            assert module is not None, f"sub_modules has no {module_key!r}"
            model_cls = getattr(importlib.import_module(mod_name), model_name)
            sub_modules[module_key] = model_cls.from_diffusers(module, device=device)
            if b200_vae and sub_modules.get("vae") is not None:
                # row f-4: the inherited post_inference keeps calling self.vae.decode(...), which
                # now runs on the B200 kernels (sduss_b200.vae.B200VAEProxy)
                from .vae import B200VAEProxy
                sub_modules["vae"] = B200VAEProxy(sub_modules["vae"], device=device)
            if b200_text_encoders:
                # row f-4, prepare half: the inherited prepare_inference -> encode_prompt keeps calling
                # self.text_encoder*(ids, output_hidden_states=True); those now run on the B200 kernels
                from .text_encoders import B200CLIPProxy, B200T5Proxy
                for key in ("text_encoder", "text_encoder_2"):
                    if sub_modules.get(key) is not None:
                        sub_modules[key] = B200CLIPProxy(sub_modules[key], device=device)
                if sub_modules.get("text_encoder_3") is not None:
                    sub_modules["text_encoder_3"] = B200T5Proxy(sub_modules["text_encoder_3"], device=device)
            return cls(**sub_modules)

        def _b200_step(self):
            step = self.__dict__.get("_b200_step_impl")
            if step is None:
                import importlib
                step_cls = getattr(importlib.import_module("sduss_b200.pipelines"), step_name)
                step = self.__dict__["_b200_step_impl"] = step_cls(getattr(self, module_key), self.scheduler)
            return step

        def denoising_step(self, *args, **kwargs):
            return self._b200_step().denoising_step(*args, **kwargs)

    B200Pipeline.__name__ = B200Pipeline.__qualname__ = "B200" + reference_cls.__name__
    return B200Pipeline

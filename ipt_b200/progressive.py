"""Headless progressive rendering — the caller side of the reference's interactive loop:

* src/main.cpp:258-269   four threads call render_sample(scene, *gui, stats) forever; every call adds one sample per pixel
* src/gui.cpp:105-137    arrow keys orbit / dolly the camera, then resetImage() + updateDisplay()
* src/gui.cpp:152-160    resetImage(): image and counters back to zero
* src/gui.cpp:83-87      updateDisplay(): normalize(glare(image, glare_cutoff))
* src/main.cpp:287-289   at exit: finalize() and save("result.png")

plus what the reference lacks: accumulator checkpoints, so a session can be stopped and resumed (the Philox counters carry
the pass index, so a resumed session produces the image of the uninterrupted one). Host glue only: every pixel operation
is an ipt_b200 C-ABI call that runs on the GPU."""
from __future__ import annotations

import copy
import os

from . import capi, checkpoint

KEY_LEFT, KEY_RIGHT, KEY_DOWN, KEY_UP = 0, 1, 2, 3


class ProgressiveSession:
    def __init__(self, scene_name: str = "box", width: int = 640, height: int = 640, passes_per_call: int = 4, seed: int = 0,
                 depth_max: int = 4, schedule=(16, 8, 4, 2), glare_cutoff: float = 1.01, device: int = 0, checkpoint_path=None):
        self.scene_name = scene_name
        self.description = capi.SceneDescription(scene_name)
        self.scene = capi.Scene(self.description, device)
        self.plane = capi.Plane(self.scene, width, height)
        self.camera = copy.copy(self.description.desc.camera)
        self.width, self.height, self.passes_per_call, self.seed = width, height, passes_per_call, seed
        self.depth_max, self.schedule, self.glare_cutoff = depth_max, list(schedule), glare_cutoff  # gui.h:24 default cutoff
        self.checkpoint_path = str(checkpoint.normalise(checkpoint_path)) if checkpoint_path else None
        self.next_pass = 0      # passes accumulated in the plane for the CURRENT camera start at `first_pass`
        self.first_pass = 0
        self.rays = 0
        if self.checkpoint_path and os.path.exists(self.checkpoint_path):
            self.resume(self.checkpoint_path)

    def identity(self) -> dict:
        """What a checkpoint must agree on to be resumed by this session (the estimator and its frame)."""
        return dict(scene=self.scene_name, depth_max=self.depth_max, schedule=list(self.schedule), plane_mode=capi.PLANE_GUI,
                    passes_per_call=self.passes_per_call)

    # -- main.cpp:262-268: one iteration of thread_func (passes_per_call samples per pixel instead of one)
    def step(self, calls: int = 1):
        for _ in range(calls):
            p = capi.default_params(width=self.width, height=self.height, depth_max=self.depth_max, schedule=self.schedule,
                                    seed=self.seed, pass_begin=self.next_pass, pass_count=self.passes_per_call, plane_mode=capi.PLANE_GUI)
            st = self.plane.render(p)
            self.next_pass += self.passes_per_call
            self.rays += st.rays
        return self.samples_per_pixel

    @property
    def samples_per_pixel(self) -> int:
        return self.next_pass - self.first_pass

    # -- gui.cpp:105-137: an arrow key moves the camera and restarts the image
    def key(self, key: int):
        capi.camera_orbit(self.camera, key)
        self.scene.set_camera(self.camera)
        self.reset_image()

    def reset_image(self):
        """Gui::resetImage (gui.cpp:152-160). Pass numbers keep growing, so the new image uses fresh random streams."""
        self.plane.clear()
        self.first_pass = self.next_pass

    # -- gui.cpp:83-87 / 186-194
    def display(self):
        """The image Gui::updateDisplay shows: normalize(glare(image, glare_cutoff)), float32 in [0, 1]."""
        return self.plane.display(self.glare_cutoff)[0]

    def set_glare_cutoff(self, wheel: int):
        """Mouse wheel (gui.cpp:141-145): `glare_cutoff *= pow(sqrt(2.0f), wheel)` — float sqrt, pow and product in double
        (std::pow(float, int) promotes), rounded back to the float member."""
        import numpy as np
        factor = np.power(np.float64(np.sqrt(np.float32(2.0))), int(wheel))
        self.glare_cutoff = float(np.float32(np.float64(np.float32(self.glare_cutoff)) * factor))

    def save(self, path="result.png"):
        """Gui::save (gui.cpp:192-194)."""
        self.plane.save_png(path)

    # -- checkpoints (no reference counterpart)
    def checkpoint(self, path=None):
        path = path or self.checkpoint_path
        cam = [list(self.camera.position), list(self.camera.direction), list(self.camera.right), list(self.camera.up)]
        return checkpoint.save(path, self.plane, self.next_pass, self.seed, identity=self.identity(), first_pass=self.first_pass,
                               camera=cam, rays=self.rays, glare_cutoff=self.glare_cutoff)

    def resume(self, path=None):
        meta = checkpoint.load(path or self.checkpoint_path, self.plane, identity=self.identity())
        self.next_pass, self.first_pass = int(meta["next_pass"]), int(meta["first_pass"])
        self.seed, self.rays, self.glare_cutoff = int(meta["seed"]), int(meta["rays"]), float(meta["glare_cutoff"])
        cam = meta["camera"]
        for name, v in zip(("position", "direction", "right", "up"), cam):
            getattr(self.camera, name)[:] = [float(x) for x in v]
        self.scene.set_camera(self.camera)

    def close(self):
        self.plane.close(); self.scene.close()

"""The body of the reference trainer's inner loop, restated so that tests can run it against vfnerf_b200 on machines where
/root/reference does not exist (train/vector_field_nerf_train.py:169-260, statement for statement; the tqdm / logging /
averaging around it is left out).  TEST INFRASTRUCTURE.  ``functions`` and ``loss`` are whatever modules the caller
plugs in: the package's (vfnerf_b200.functions, vfnerf_b200.losses.VFLoss) on the GPU."""
import torch


def train_step(trainer, train_data, epoch):
    """One iteration of VectorFieldNerfTrainer.train_epoch.  ``trainer`` has .model, .loss, .dataset, .config and
    .functions like the reference object (self.model / self.loss / self.dataset / self.config, module `functions`)."""
    self, functions = trainer, trainer.functions
    dev = self.config.vf_nerf_config.cuda_config.device
    # :172-174
    pixels = train_data["uv"].squeeze(0).to(dev)
    intrinsics = train_data["intrinsics"].squeeze(0).to(dev)
    pose = train_data["pose"].squeeze(0).to(dev)
    # :177
    outputs = self.model.render(pose, pixels, intrinsics, epoch, self.dataset.white_bkgd)
    n_extra = (outputs.points_coarse.shape[0] * outputs.points_coarse.shape[1]) // 10
    # :180-195
    if self.dataset.get_vf_init_method()[0] == "center" and self.config.dataset_config.dataset_name != "deepfashion":
        supervised_normals, gt_normals = functions.get_border_indices_and_gt(
            outputs.points_coarse, outputs.coarse_normals, self.dataset.get_bounds()[1],
            self.config.dataset_config.border_radius, self.dataset.get_centroid(dev))
        border_points, border_gt_normals = functions.sample_border_points(
            self.dataset.get_bounds()[1] / 2 - self.config.dataset_config.border_radius, self.dataset.get_bounds()[1] / 2,
            n_extra, self.dataset.get_centroid(dev), outputs.points_coarse.device)
        border_normals = self.model.vector_field_network(border_points)[:, :3]
        supervised_normals = torch.cat([supervised_normals, border_normals], dim=0)
        gt_normals = torch.cat([gt_normals, border_gt_normals], dim=0)
    else:
        # :196-218
        supervised_normals = torch.empty(0, 3).to(dev)
        gt_normals = torch.empty(0).to(dev)
        if self.config.vf_nerf_config.border_supervision:
            border_points, border_gt_normals = functions.sample_border_points(
                self.dataset.get_bounds()[1] - 5 * self.config.dataset_config.border_radius, self.dataset.get_bounds()[1],
                n_extra, self.dataset.get_centroid(dev), outputs.points_coarse.device)
            supervised_normals = torch.cat([supervised_normals, self.model.vector_field_network(border_points)[:, :3]], dim=0)
            gt_normals = torch.cat([gt_normals, border_gt_normals], dim=0)
        if self.config.vf_nerf_config.center_supervision:
            ray_center_normals, ray_center_gt_normals = functions.get_center_indices_and_gt(
                outputs.points_coarse, outputs.coarse_normals, self.dataset.get_centroid(dev),
                self.config.dataset_config.border_radius)
            center_points, center_gt_normals = functions.sample_center_points(
                self.dataset.get_centroid(dev), self.config.dataset_config.border_radius, n_extra,
                outputs.points_coarse.device)
            supervised_normals = torch.cat([supervised_normals, ray_center_normals,
                                            self.model.vector_field_network(center_points)[:, :3]], dim=0)
            gt_normals = torch.cat([gt_normals, ray_center_gt_normals, center_gt_normals], dim=0)
    # :221-232
    predictions = {"rgb": outputs.coarse_rgb_values, "depth": outputs.coarse_depth_map,
                   "normals": outputs.coarse_normals.reshape(-1, 3), "supervised_normals": supervised_normals,
                   "directional_derivatives": outputs.directional_derivtives}
    ground_truth = {"rgb": train_data["rgb"].reshape(-1, 3).to(dev), "depth": train_data["depth"].squeeze(0).to(dev),
                    "supervised_normals": gt_normals}
    # :234 (the fine loss of :237-250 never runs: outputs.fine_normals is None, vector_field_nerf.py:279-283)
    loss, losses_dict = self.loss(predictions, ground_truth, epoch)
    total_loss = loss
    # :252-260
    self.model.optimizer.zero_grad()
    total_loss.backward()
    torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.config.vf_nerf_config.scheduler_config.clip_norm)
    self.model.optimizer.step()
    self.model.scheduler.step()
    return total_loss.detach(), losses_dict, outputs

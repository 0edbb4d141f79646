"""CoFusion::processFrame sequencing (Core/CoFusion.cpp:171-545) driven through the CPU oracle --
the checker for cofusion_b200.CoFusion.  TEST INFRASTRUCTURE."""
import numpy as np

import orc


class OraclePipeline:
    def __init__(self, W, H, K, max_surfels=1 << 20, conf_global=10.0, depth_cutoff=5.0, max_depth=20.0,
                 icp_weight=10.0, time_delta=200, outlier_coeff=3.0):
        self.W, self.H, self.K = W, H, K
        self.map = orc.OrcMap(W, H, K, max_surfels)
        self.odom = orc.OrcOdometry(W, H, K)
        self.pose = np.eye(4, dtype=np.float32)
        self.last_pose = np.eye(4, dtype=np.float32)
        self.tick = 1
        self.conf = conf_global
        self.depth_cutoff, self.max_depth = depth_cutoff, max_depth
        self.icp_weight, self.time_delta, self.outlier = icp_weight, time_delta, outlier_coeff
        self.stats = None

    def predict(self, rgb, df):
        self.map.combined_predict(self.pose, self.max_depth, self.conf, self.tick, self.tick, self.time_delta)
        self.map.fill_in(rgb, df, False, False)

    def process_frame(self, rgb, depth, mask=None):
        mask = np.zeros((self.H, self.W), np.uint8) if mask is None else mask
        df = orc.bilateral(depth, self.depth_cutoff)
        if self.tick == 1:
            self.map.initialise(rgb, depth, df, self.tick, self.max_depth)
            self.odom.init_first_rgb(rgb)
        else:
            m = self.map
            if m.requires_fill_in():  # Model::initICP with doFillIn (Model.cpp:354-356)
                v4, n4, img = m.view(9), m.view(10), m.view(8)
            else:
                v4, n4, img = m.view(5), m.view(6), m.view(4)
            self.last_pose = self.pose.copy()
            self.odom.init_model(v4, n4, img, self.pose)
            self.odom.init_frame(df, rgb, self.max_depth)
            self.pose, self.stats, _, _ = self.odom.track(self.pose, icp_weight=self.icp_weight)
            self.predict(rgb, df)
            w = orc.OrcMap.fusion_weight(self.pose, self.last_pose, 1.0)
            m.predict_indices(self.pose, self.tick, self.max_depth, self.time_delta)
            m.fuse(self.pose, self.tick, rgb, mask, depth, df, self.max_depth, w, 0)
            m.predict_indices(self.pose, self.tick, self.max_depth, self.time_delta)
            m.clean(self.pose, self.tick, self.conf, self.time_delta, df, mask, 0, self.outlier)
        self.predict(rgb, df)
        self.tick += 1

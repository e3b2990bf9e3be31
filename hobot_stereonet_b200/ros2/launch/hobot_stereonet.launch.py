"""Launch file of the B200 stereonet node: the reference's launch arguments (config_file, model_file;
stereonet_infer/launch/hobot_stereonet.launch.py:35-40) plus the geometry an .hbm used to compile in.
`model_file` names a SNB2WGT1 weight blob instead of an .hbm."""
import os

from ament_index_python.packages import get_package_prefix
from launch import LaunchDescription
from launch.actions import DeclareLaunchArgument
from launch.substitutions import LaunchConfiguration
from launch_ros.actions import Node


def generate_launch_description():
    pkg_path = os.path.join(get_package_prefix("hobot_stereonet"), "lib/hobot_stereonet")
    args = [
        DeclareLaunchArgument("config_file", default_value=pkg_path + "/config/hobot_stereonet_config.json"),
        DeclareLaunchArgument("model_file", default_value=pkg_path + "/config/hobot_stereonet.snb"),
        DeclareLaunchArgument("sub_hbmem_topic_name", default_value="hbmem_stereo_img"),
        DeclareLaunchArgument("ros_img_topic_name", default_value="/stereonet_node_output"),
        DeclareLaunchArgument("model_in_h", default_value="720"),
        DeclareLaunchArgument("model_in_w", default_value="1280"),
        DeclareLaunchArgument("K", default_value="4"),
        DeclareLaunchArgument("D", default_value="12"),
        DeclareLaunchArgument("device", default_value="0"),
        DeclareLaunchArgument("precision", default_value="tc"),
    ]
    names = ["config_file", "model_file", "sub_hbmem_topic_name", "ros_img_topic_name", "model_in_h", "model_in_w", "K",
             "D", "device", "precision"]
    node = Node(
        package="hobot_stereonet",
        executable="hobot_stereonet",
        output="screen",
        parameters=[{n: LaunchConfiguration(n)} for n in names],
        arguments=["--ros-args", "--log-level", "warn"],
    )
    return LaunchDescription(args + [node])

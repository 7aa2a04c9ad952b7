"""B200-native batched RBP trajectory-QP engine: drop-in for the inside of SwarmPlanning::RBPPlanner::update()
(/root/reference/swarm_planner/include/rbp_planner.hpp L33-L84). See DESIGN.md."""

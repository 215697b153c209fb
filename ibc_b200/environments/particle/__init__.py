from ibc_b200.environments.particle.particle import ParticleEnv, HistoryWrapper
from ibc_b200.environments.particle.particle_oracles import ParticleOracle

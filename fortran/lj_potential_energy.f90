!******************************************************************************************************
! Drop-in replacement for scripts/physics/lj_potential_energy.f90: same module name, same
! procedure name and dummy list (lj_potential_energy.f90:46-52); the body only unpacks the
! derived types and calls the CUDA library through ISO_C_BINDING.
!******************************************************************************************************
module lj_potential_energy
  use define_precision, only: dp_kind, int_kind
  use md_types,         only: sim_params, sim_state, pi
  use ljmd_c_api
  implicit none

  logical, parameter :: use_tail_corrections = .true.   ! kept for source compatibility (:36)

contains

  subroutine compute_lj_potential_energy(params, state, epot, d_epot, dd_epot)
    type(sim_params), intent(in)    :: params
    type(sim_state),  intent(inout) :: state
    real(kind=dp_kind), intent(out) :: epot, d_epot, dd_epot
    type(ljmd_params) :: cp

    if (.not. allocated(state%rx)) stop 'compute_lj_potential_energy(): state arrays are not allocated.'
    cp = to_c_params(params%n, params%num_cells, params%box_length, params%inv_box_length, &
                     params%volume, params%density, params%dt, params%dt_half, params%dt_square_half, &
                     params%rc, params%rc_square, params%length_scale, params%inv_length_scale)
    ! contiguous allocatables are passed by base address (assumed-size dummies, no descriptor)
    call ljmd_stop_on_error(ljmd_compute_lj(cp, state%rx, state%ry, state%rz, &
                                            state%ax, state%ay, state%az, epot, d_epot, dd_epot))
  end subroutine compute_lj_potential_energy

end module lj_potential_energy
